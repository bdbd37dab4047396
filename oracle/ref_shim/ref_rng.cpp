// oracle/ref_shim/ref_rng.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Link-time seam for the reference's random numbers.  Upstream declares
//     uint RandomUInt(); float RandomFloat();            (template/precomp.h:344-349)
// and defines them over ONE global xorshift32 state (template/template.cpp:666-685), which
// makes every pixel depend on every earlier pixel.  The oracle keeps the generator
// (Marsaglia xorshift32, shifts 13/17/5, float = uint * 2.3283064365387e-10f) but gives
// every (pixel, sample) path its own start state, so the CPU reference and the GPU
// wavefront consume identical streams whatever the scheduling:
//
//     state = WangHash( WangHash((y*W + x + 1) * 17) + s ),  0 -> 1
//
// i.e. the template's own per-thread recipe "seed using WangHash((threadidx+1)*17)"
// (cl/tools.cl:1-2) applied per pixel and re-hashed with the sample index.
// Also hosts the few template.cpp definitions the hot path links against.
#include "precomp.h"

#define STB_IMAGE_IMPLEMENTATION
#include "lib/stb_image.h"

static thread_local uint t_state = 0x12345678u;   // upstream's default seed (template.cpp:667)

static inline uint wang_hash(uint s) {
	s = (s ^ 61u) ^ (s >> 16);
	s *= 9u;
	s = s ^ (s >> 4);
	s *= 0x27d4eb2du;
	s = s ^ (s >> 15);
	return s;
}

extern "C" unsigned agpt_ref_stream_seed(unsigned pixel_index, unsigned sample) {
	uint s = wang_hash(wang_hash((pixel_index + 1u) * 17u) + sample);
	return s ? s : 1u;
}
extern "C" void agpt_ref_seed_path(unsigned pixel_index, unsigned sample) {
	t_state = agpt_ref_stream_seed(pixel_index, sample);
}
extern "C" void agpt_ref_set_state(unsigned s) { t_state = s ? s : 1u; }
extern "C" unsigned agpt_ref_get_state() { return t_state; }

uint RandomUInt() {
	t_state ^= t_state << 13;
	t_state ^= t_state >> 17;
	t_state ^= t_state << 5;
	return t_state;
}
float RandomFloat() { return RandomUInt() * 2.3283064365387e-10f; }
float Rand(float range) { return RandomFloat() * range; }
uint RandomUInt(uint& s) {
	s ^= s << 13;
	s ^= s >> 17;
	s ^= s << 5;
	return s;
}
float RandomFloat(uint& s) { return RandomUInt(s) * 2.3283064365387e-10f; }

// 4x4 row-major product; only reached when a scene composes transforms (myapp.cpp:22).
mat4 operator*(const mat4& a, const mat4& b) {
	mat4 r;
	for (int row = 0; row < 4; row++)
		for (int col = 0; col < 4; col++) {
			float acc = 0;
			for (int k = 0; k < 4; k++) acc += a.cell[row * 4 + k] * b.cell[k * 4 + col];
			r.cell[row * 4 + col] = acc;
		}
	return r;
}
