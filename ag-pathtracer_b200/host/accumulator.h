// host/accumulator.h -- Accumulator with the reference's surface (/root/reference/myapp.h:8-68):
// a W x H buffer of float3 (16 bytes each, i.e. float4) summed over samples, row 0 = top of
// the image (AddSample stores to row height-1-y).  On the host it is the landing buffer for
// the device accumulator; the per-sample adds happen on the GPU.
#pragma once

#include "precomp.h"
#include "agpt.h"

namespace Tmpl8 {

class Accumulator {
public:
	Accumulator(int w, int h, const int2& scrPos = int2(0, 0)) : width(w), height(h), screenPos(scrPos), samples(0) {
		// page-locked when a CUDA device is there (film copies at PCIe speed), plain aligned memory otherwise
		void* p = nullptr;
		size_t bytes = ((size_t)w * h * sizeof(float3) + 63) / 64 * 64;
		pinned = agpt_host_alloc(bytes, &p) == AGPT_OK;
		pixels = (float3*)(pinned ? p : aligned_alloc(64, bytes));
		Clear();
	}
	// wrap caller-owned film memory (not freed here)
	Accumulator(int w, int h, float3* external) : width(w), height(h), screenPos(0, 0), pixels(external), samples(0), pinned(false), owned(false) {}
	~Accumulator() { if (owned) { if (pinned) agpt_host_free(pixels); else free(pixels); } }
	Accumulator(const Accumulator&) = delete;
	Accumulator& operator=(const Accumulator&) = delete;

	inline void AddSample(int x, int y, const float3& clr) { pixels[(height - 1 - y) * width + x] += clr; }
	inline void IncrementSampleCount() { samples++; }
	inline int NumSamples() const { return samples; }
	inline void Clear() {
		memset((void*)pixels, 0, (size_t)width * height * sizeof(float3));
		samples = 0;
	}
	inline float2 PixelToFilm(const float2& p) const { return float2(p.x / width, p.y / height); }

	// device hand-off
	float3* Pixels() { return pixels; }
	const float3* Pixels() const { return pixels; }
	void SetNumSamples(int s) { samples = s; }

	const int width, height;

private:
	int2 screenPos;
	float3* pixels;
	int samples;
	bool pinned = false, owned = true;
};

} // namespace Tmpl8
using namespace Tmpl8;
