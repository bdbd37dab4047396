// host/lights.h -- Light, UniformInfiniteLight, AreaLight with the reference's
// construction surface (/root/reference/lights.h:23-87).
//
// Host lights are descriptors.  Sample_Li / Pdf_Li / Le (lights.cpp:10-28,115-130) and the
// visibility test (lights.cpp:10-12) are evaluated on the device by the shade and any-hit
// kernels.  InfiniteAreaLight (HDR environment map, lights.cpp:31-112) is the next row of
// SURVEY 8f; its asset is not in the reference repo.
#pragma once

#include "precomp.h"
#include "intersectable.h"

class Light {
public:
	virtual ~Light() {}
	virtual bool IsInfinite() const { return false; }
	virtual int Kind() const = 0;      // AGPT_LIGHT_*
	virtual float3 Emission() const = 0;
};

class UniformInfiniteLight : public Light {
public:
	UniformInfiniteLight(const float3& l) : Lemit(l) {}
	bool IsInfinite() const override { return true; }
	int Kind() const override { return AGPT_LIGHT_UNIFORM_INFINITE; }
	float3 Emission() const override { return Lemit; }
protected:
	const float3 Lemit;
};

// Wraps any Intersectable; as upstream, only Sphere can be sampled (others emit when hit
// by camera / specular rays only).  Emission is two-sided (lights.h:82).
class AreaLight : public Light {
public:
	AreaLight(shared_ptr<Intersectable> shape, const float3& l) : Shape(shape), Lemit(l) { shape->SetAreaLight(this); }
	int Kind() const override { return AGPT_LIGHT_AREA; }
	float3 Emission() const override { return Lemit; }
	shared_ptr<Intersectable> Shape;
protected:
	const float3 Lemit;
};
