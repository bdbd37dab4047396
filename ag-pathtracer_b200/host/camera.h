// host/camera.h -- Ray, CameraDesc and Camera with the reference's public surface
// (/root/reference/camera.h:3-15 Ray, :17-25 CameraDesc, :27-106 Camera).
//
// Host side of the drop-in: the constructor derives the camera basis with the reference's
// arithmetic (tan of the half field of view, focus distance = |lookat - lookfrom|,
// camera.h:38-56,77-90) and Export() ships the six vectors + lens radius to the device
// (agpt_camera).  Ray generation itself (Camera::GetRay, camera.h:58-64) runs on the GPU,
// inside the path-generation kernel; there is no host GetRay.
#pragma once

#include "precomp.h"
#include "agpt.h"

class Ray {
public:
	Ray() = default;
	Ray(float3 o, float3 d, float t = FLT_MAX) : O(o), D(normalize(d)), t(t) {}
	float3 at(float s) const { return O + s * D; }
	float3 O, D;
	mutable float t;
};

struct CameraDesc {
	float3 lookfrom;
	float3 lookat;
	float3 vup;
	float aspect_ratio;
	float vfov = 45;
	float focus_dist = 1.0f;   // ignored upstream too: focus distance is |lookat - lookfrom| (camera.h:78)
	float aperture = 0.0f;
};

class Camera {
public:
	Camera(const CameraDesc& d) : Camera(d.lookfrom, d.lookat, d.vup, d.aspect_ratio, d.vfov, d.aperture) {}

	Camera(float3 lookfrom, float3 lookat, float3 vup, float aspect_ratio, float vfov, float aperture)
		: lookat(lookat), vup(vup) {
		float halfHeight = std::tan(radians(vfov) / 2);
		viewport_height = 2 * halfHeight;
		viewport_width = aspect_ratio * viewport_height;
		lens_radius = aperture / 2;
		Place(lookfrom);
	}

	float3 GetOrigin() const { return origin; }

	// The derived state the device needs (agpt_set_camera).
	agpt_camera Export() const {
		agpt_camera c;
		auto put = [](float* dst, const float3& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; };
		put(c.origin, origin); put(c.lower_left_corner, lower_left_corner);
		put(c.horizontal, horizontal); put(c.vertical, vertical);
		put(c.u, u); put(c.v, v);
		c.lens_radius = lens_radius;
		return c;
	}

protected:
	// camera.h:77-90: w looks away from the scene, image plane sits at the focus distance.
	void Place(float3 lookfrom) {
		focus_dist = length(lookat - lookfrom);
		w = normalize(lookfrom - lookat);
		u = normalize(cross(vup, w));
		v = cross(w, u);
		origin = lookfrom;
		horizontal = focus_dist * viewport_width * u;
		vertical = focus_dist * viewport_height * v;
		lower_left_corner = origin - horizontal / 2 - vertical / 2 - focus_dist * w;
	}

	float3 lookat, vup;
	float focus_dist, lens_radius, viewport_width, viewport_height;
	float3 origin, u, v, w;
	float3 lower_left_corner, horizontal, vertical;
};
