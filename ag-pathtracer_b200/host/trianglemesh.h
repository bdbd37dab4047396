// host/trianglemesh.h -- index_type and TriangleMesh with the reference's construction
// surface (/root/reference/trianglemesh.h:3-56; CreateBackdrop trianglemesh.cpp:232-318).
//
// A host mesh is indexed geometry only.  Ray/triangle intersection (Moller-Trumbore,
// trianglemesh.cpp:7-43,117-155) and the shading-frame reconstruction (:45-113) are device
// code; ExportTriangles() lays the triangles out in the order the device reads them.
// LoadObj is out of scope (no .obj assets).
#pragma once

#include "precomp.h"
#include "intersectable.h"

struct index_type {
	int vertex_index, normal_index, texcoord_index;
	index_type(int idx) : vertex_index(idx), normal_index(idx), texcoord_index(idx) {}
	index_type(int v, int n, int t) : vertex_index(v), normal_index(n), texcoord_index(t) {}
};

// Leaf-ordered triangle arrays in the device layout (agpt_mesh_desc).
struct FlatTriangles {
	std::vector<float> verts;     // 3 x float4 per triangle
	std::vector<int32_t> ids;     // original triangle number
	std::vector<float> normals;   // 3 x float4 per triangle, empty if the mesh has none
	std::vector<float> uvs;       // 3 x float2 per triangle, empty if the mesh has none
};

class TriangleMesh : public Intersectable {
public:
	// Like upstream, construction MOVES the four arrays out of the caller's vectors
	// (trianglemesh.h:16-21) -- and out of the source mesh in the shared_ptr overload.
	TriangleMesh(vector<index_type>& indices, vector<float3>& vertices, vector<float3>& normals,
			vector<float2>& texcoords, shared_ptr<Material> mat)
		: Intersectable(mat), vertices(std::move(vertices)), normals(std::move(normals)),
		  texcoords(std::move(texcoords)), indices(std::move(indices)) {}
	TriangleMesh(shared_ptr<TriangleMesh> src, shared_ptr<Material> mat)
		: TriangleMesh(src->indices, src->vertices, src->normals, src->texcoords, mat) {}

	int Kind() const override { return AGPT_PRIM_MESH; }
	int NumTriangles() const { return (int)indices.size() / 3; }

	// order[j] = original triangle number stored in slot j (identity for a plain mesh).
	FlatTriangles ExportTriangles(const std::vector<int32_t>& order) const {
		FlatTriangles f;
		size_t n = order.size();
		f.ids = order;
		f.verts.assign(n * 12, 0.f);
		if (!normals.empty()) f.normals.assign(n * 12, 0.f);
		if (!texcoords.empty()) f.uvs.assign(n * 6, 0.f);
		for (size_t j = 0; j < n; j++)
			for (int k = 0; k < 3; k++) {
				const index_type& ix = indices[3 * (size_t)order[j] + k];
				const float3& v = vertices[ix.vertex_index];
				float* dst = &f.verts[j * 12 + 4 * k];
				dst[0] = v.x; dst[1] = v.y; dst[2] = v.z;
				if (!normals.empty()) {
					const float3& nn = normals[ix.normal_index];
					float* nd = &f.normals[j * 12 + 4 * k];
					nd[0] = nn.x; nd[1] = nn.y; nd[2] = nn.z;
				}
				if (!texcoords.empty()) {
					const float2& t = texcoords[ix.texcoord_index];
					f.uvs[j * 6 + 2 * k] = t.x; f.uvs[j * 6 + 2 * k + 1] = t.y;
				}
			}
		return f;
	}

	static std::shared_ptr<TriangleMesh> CreateBackdrop(const float3& origin, const float3& size, float radius,
			int steps, std::shared_ptr<Material> material);

protected:
	vector<float3> vertices;
	vector<float3> normals;
	vector<float2> texcoords;
	vector<index_type> indices;
};

// Photo-studio backdrop: back wall, quarter-circle bevel of `steps` segments, floor; a strip
// of vertex pairs (+width/2, -width/2) joined by two triangles per segment.  Same vertices,
// normals, uvs and index order as upstream (trianglemesh.cpp:232-318) so the BVH built over
// it is identical.
inline std::shared_ptr<TriangleMesh> TriangleMesh::CreateBackdrop(const float3& origin, const float3& size,
		float radius, int steps, std::shared_ptr<Material> material) {
	const float halfW = size[0] / 2, height = size[1], depth = size[2];
	std::vector<float3> vertices, normals;
	std::vector<float2> texcoords;
	float row = 0;   // texture u advances by one per strip row
	auto addRow = [&](float y, float z, const float3& n) {
		vertices.push_back(origin + float3(halfW, y, z));
		vertices.push_back(origin + float3(-halfW, y, z));
		normals.push_back(n); normals.push_back(n);
		texcoords.push_back({ row, 0 }); texcoords.push_back({ row, 1 });
		row += 1;
	};
	const float3 wallN(0, 0, -1), floorN(0, 1, 0);
	addRow(height, 0, wallN);
	addRow(radius * 1.1f, 0, wallN);          // guard quad: wall does not share normals with the bevel
	auto stepAngle = PI / (2 * steps);
	for (auto i = 0; i <= steps; i++) {
		auto zRot = std::cos(stepAngle * i);
		auto yRot = -std::sin(stepAngle * i);
		addRow(yRot * radius + radius, zRot * radius - radius, normalize(float3(0, -yRot, -zRot)));
	}
	addRow(0, -radius * 1.1f, floorN);        // guard quad before the floor
	addRow(0, -depth, floorN);

	std::vector<index_type> indices;
	int parts = 4 + steps;
	for (int i = 0; i < parts; i++) {
		int a = 2 * i, b = 2 * i + 1, c = 2 * (i + 1), d = 2 * (i + 1) + 1;
		for (int v : { a, c, b, c, d, b }) indices.push_back(index_type(v));
	}
	return make_shared<TriangleMesh>(indices, vertices, normals, texcoords, material);
}
