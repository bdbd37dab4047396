// host/intersectable.h -- Intersectable, Sphere, Plane with the reference's construction
// surface (/root/reference/intersectable.h:17-61 Intersectable, :119-157 Plane, :159-322 Sphere).
//
// Host objects only describe shapes.  The virtuals the reference calls per ray
// (Intersect / IntersectP / Sample / Pdf) have no host implementation here: closest-hit,
// any-hit and light sampling for these shapes are __device__ code in
// ag-pathtracer_b200/csrc/ (sphere: intersectable.h:164-226,239-317; plane: :123-150).
#pragma once

#include "precomp.h"
#include "material.h"
#include "camera.h"

class AreaLight;

class Intersectable {
public:
	Intersectable(shared_ptr<Material> material) : material(material), arealight(nullptr) {}
	virtual ~Intersectable() {}

	virtual int Kind() const = 0;          // AGPT_PRIM_* row type in the device primitive table
	virtual float Area() const { return 0.f; }

	const AreaLight* GetAreaLight() const { return arealight; }
	void SetAreaLight(const AreaLight* light) { arealight = light; }
	const Material* GetMaterial() const { return material.get(); }

private:
	std::shared_ptr<Material> material;
	const AreaLight* arealight;
};

// XZ rectangle through O with normal +y; `size` is full width/depth (intersectable.h:121).
class Plane : public Intersectable {
public:
	Plane(float3 o, float2 size, shared_ptr<Material> m) : Intersectable(m), O(o), HalfSize(size / 2) {}
	int Kind() const override { return AGPT_PRIM_PLANE; }
	float Area() const override { return 4.f * HalfSize.x * HalfSize.y; }
	agpt_plane Export() const {
		agpt_plane p;
		memset(&p, 0, sizeof(p));
		p.o[0] = O.x; p.o[1] = O.y; p.o[2] = O.z;
		p.half_x = HalfSize.x; p.half_z = HalfSize.y;
		return p;
	}
	float3 O;
	float2 HalfSize;
};

class Sphere : public Intersectable {
public:
	Sphere(float3 center, float radius, shared_ptr<Material> m)
		: Intersectable(m), Center(center), r(radius), r2(radius * radius) {}
	int Kind() const override { return AGPT_PRIM_SPHERE; }
	float Area() const override { return 4.f * PI * r2; }
	agpt_sphere Export() const {
		agpt_sphere s;
		memset(&s, 0, sizeof(s));
		s.center[0] = Center.x; s.center[1] = Center.y; s.center[2] = Center.z;
		s.r = r; s.r2 = r2;
		return s;
	}
	float3 Center;
	float r, r2;
};
