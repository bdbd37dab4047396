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

	void Build(int maxPrims) {
		const int nTris = NumTriangles();
		prims.reserve(nTris);
		for (int t = 0; t < nTris; t++) {
			BuildPrim p;
			p.tri = t;
			for (int k = 0; k < 3; k++) p.box.Grow(&vertices[indices[3 * t + k].vertex_index].x);
			for (int a = 0; a < 3; a++) p.centroid[a] = (p.box.lo[a] + p.box.hi[a]) * 0.5f;
			prims.push_back(p);
		}
		nodes.assign(2, BVHNode{});   // root + the unused slot 1
		if (nTris > 0) BuildRange(0, 0, nTris, maxPrims);
		leafOrder.resize(nTris);
		for (int j = 0; j < nTris; j++) leafOrder[j] = prims[j].tri;
		prims.clear();
		prims.shrink_to_fit();
	}

	void StoreNode(int slot, const Box& b, int first, int count) {
		BVHNode& n = nodes[slot];
		for (int a = 0; a < 3; a++) { n.bmin3[a] = b.lo[a]; n.bmax3[a] = b.hi[a]; }
		n.first = first;
		n.count = count;
	}

	// Decide node `slot` for prims[start,end): leaf, or split at `mid` and recurse.
	void BuildRange(int slot, int start, int end, int maxPrims) {
		Box bounds;
		for (int i = start; i < end; i++) bounds.Grow(prims[i].box);
		const int n = end - start;
		if (n == 1) { StoreNode(slot, bounds, start, n); return; }

		Box centroidBox;
		for (int i = start; i < end; i++) centroidBox.Grow(prims[i].centroid);
		const int axis = centroidBox.LongestAxis();
		if (centroidBox.lo[axis] == centroidBox.hi[axis]) { StoreNode(slot, bounds, start, n); return; }

		int mid = (start + end) / 2;
		if (n <= 2) {
			std::nth_element(prims.begin() + start, prims.begin() + mid, prims.begin() + end,
				[axis](const BuildPrim& a, const BuildPrim& b) { return a.centroid[axis] < b.centroid[axis]; });
		}
		else {
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
			if (n > maxPrims || minCost < leafCost) {
				BuildPrim* pmid = std::partition(&prims[start], &prims[end - 1] + 1,
					[&](const BuildPrim& p) { return BucketOf(centroidBox, p, axis) <= minBucket; });
				mid = (int)(pmid - &prims[0]);
			}
			else { StoreNode(slot, bounds, start, n); return; }
		}

		int pair = (int)nodes.size();          // children go to the next free even index
		nodes.resize(nodes.size() + 2);
		StoreNode(slot, bounds, pair, 0);
		BuildRange(pair, start, mid, maxPrims);
		BuildRange(pair + 1, mid, end, maxPrims);
	}

	std::vector<BuildPrim> prims;
	std::vector<BVHNode> nodes;
	std::vector<int32_t> leafOrder;
};
