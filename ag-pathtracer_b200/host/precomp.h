// host/precomp.h -- vector math the reference's scene-building code leans on.
//
// Drop-in counterpart of the parts of /root/reference/template/precomp.h and
// template/common.h that host-side scene construction touches (vector structs
// precomp.h:124-205, operators and helpers :363-789, colour helpers common.h:21-51).
// Written from scratch for the host mirror: only what building a Scene, a Camera and a
// BVH needs.  No intersection, shading or sampling code lives on the host -- that is the
// device's job (include/agpt.h); there is no CPU fallback.
//
// Parity notes: float3 is 16 bytes like upstream (OpenCL layout, precomp.h:165-172);
// normalize(v) multiplies by 1/sqrtf(dot) (precomp.h:366,735) -- not v/len, not rsqrt --
// because Camera and scene vertices built here must match the oracle bit for bit.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

using namespace std;   // the reference leaks std into every TU (precomp.h:33); scene code relies on it

typedef unsigned char uchar;
typedef unsigned int uint;

#define PI      3.14159265358979323846264f
#define INVPI   0.31830988618379067153777f
#define INV2PI  0.15915494309189533576888f
#define TWOPI   6.28318530717958647692528f
#define EPSILON 0.0001f

struct alignas(8) float2 {
	float2() = default;
	float2(float a, float b) : x(a), y(b) {}
	float2(float a) : x(a), y(a) {}
	float x, y;
	float operator[](int n) const { return n ? y : x; }
};

struct alignas(8) int2 {
	int2() = default;
	int2(int a, int b) : x(a), y(b) {}
	int2(int a) : x(a), y(a) {}
	int x, y;
};

struct alignas(16) float3 {
	float3() = default;
	float3(float a, float b, float c) : x(a), y(b), z(c) {}
	float3(float a) : x(a), y(a), z(a) {}
	float x, y, z, dummy = 0.f;     // (upstream leaves the padding lane uninitialised; nothing may depend on it)
	float operator[](int n) const { return (&x)[n]; }
};

inline float2 make_float2(float a, float b) { return float2(a, b); }
inline float3 make_float3(float a, float b, float c) { return float3(a, b, c); }
inline float3 make_float3(float s) { return float3(s, s, s); }

inline float2 operator+(const float2& a, const float2& b) { return { a.x + b.x, a.y + b.y }; }
inline float2 operator-(const float2& a, const float2& b) { return { a.x - b.x, a.y - b.y }; }
inline float2 operator*(const float2& a, float s) { return { a.x * s, a.y * s }; }
inline float2 operator*(float s, const float2& a) { return { s * a.x, s * a.y }; }
inline float2 operator/(const float2& a, float s) { return { a.x / s, a.y / s }; }

inline float3 operator-(const float3& a) { return { -a.x, -a.y, -a.z }; }
inline float3 operator+(const float3& a, const float3& b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline float3 operator-(const float3& a, const float3& b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline float3 operator*(const float3& a, const float3& b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline float3 operator*(const float3& a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline float3 operator*(float s, const float3& a) { return { s * a.x, s * a.y, s * a.z }; }
inline float3 operator/(const float3& a, float s) { return { a.x / s, a.y / s, a.z / s }; }
inline void operator+=(float3& a, const float3& b) { a.x += b.x; a.y += b.y; a.z += b.z; }
inline void operator-=(float3& a, const float3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
inline void operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
inline void operator*=(float3& a, const float3& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
inline void operator/=(float3& a, float s) { a.x /= s; a.y /= s; a.z /= s; }

inline float sqr(float x) { return x * x; }
inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float sqrLength(const float3& v) { return dot(v, v); }
inline float length(const float3& v) { return sqrtf(dot(v, v)); }
inline float3 normalize(const float3& v) { float inv = 1.0f / sqrtf(dot(v, v)); return v * inv; }
inline float3 cross(const float3& a, const float3& b) {
	return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}
inline float3 Lerp(float t, const float3& a, const float3& b) { return (1 - t) * a + t * b; }
inline bool IsBlack(const float3& v) { return v.x == 0 && v.y == 0 && v.z == 0; }
inline float clamp(float f, float lo, float hi) { float m = f < hi ? f : hi; return lo > m ? lo : m; }
inline std::ostream& operator<<(std::ostream& os, const float3& v) { return os << "(" << v.x << ", " << v.y << ", " << v.z << ")"; }

inline float radians(float degrees) { return degrees * PI / 180.0f; }
inline float3 rgb2lin(float3 c) { return { std::pow(c.x, 2.2f), std::pow(c.y, 2.2f), std::pow(c.z, 2.2f) }; }
inline float3 hex2lin(int hex) {
	float3 rgb(((hex >> 16) & 0xFF) / 255.f, ((hex >> 8) & 0xFF) / 255.f, (hex & 0xFF) / 255.f);
	return rgb2lin(rgb);
}

// Row-major 4x4 for baking transforms into vertices at scene-build time (the reference has
// no per-primitive transforms; LoadObj bakes them, trianglemesh.cpp:157).
struct mat4 {
	float cell[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
	static mat4 Identity() { return mat4(); }
	static mat4 Translate(float x, float y, float z) { mat4 m; m.cell[3] = x; m.cell[7] = y; m.cell[11] = z; return m; }
	static mat4 Scale(float s) { mat4 m; m.cell[0] = m.cell[5] = m.cell[10] = s; return m; }
	static mat4 RotateY(float a) { mat4 m; float c = cosf(a), s = sinf(a); m.cell[0] = c; m.cell[2] = s; m.cell[8] = -s; m.cell[10] = c; return m; }
	float3 TransformPoint(const float3& v) const {
		return { cell[0] * v.x + cell[1] * v.y + cell[2] * v.z + cell[3],
			cell[4] * v.x + cell[5] * v.y + cell[6] * v.z + cell[7],
			cell[8] * v.x + cell[9] * v.y + cell[10] * v.z + cell[11] };
	}
	float3 TransformVector(const float3& v) const {
		return { cell[0] * v.x + cell[1] * v.y + cell[2] * v.z,
			cell[4] * v.x + cell[5] * v.y + cell[6] * v.z,
			cell[8] * v.x + cell[9] * v.y + cell[10] * v.z };
	}
};
inline mat4 operator*(const mat4& a, const mat4& b) {
	mat4 r;
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++) {
			float acc = 0;
			for (int k = 0; k < 4; k++) acc += a.cell[4 * i + k] * b.cell[4 * k + j];
			r.cell[4 * i + j] = acc;
		}
	return r;
}
