// host/material.h -- Material, DisneyMaterial, MirrorMaterial with the reference's
// construction surface (/root/reference/material.h:6-88).
//
// On the host a material is just its parameters: the constructor derives, with the
// reference's float arithmetic, the constants its BxDF objects would hold (diffuse/retro
// reflectance, Trowbridge-Reitz alphas, Schlick R0; material.h:14-49, disney.h:22) and
// Export() packs them into the 64-byte agpt_material record the shade kernel reads.
// BSDF evaluation and sampling run on the device only.
#pragma once

#include "precomp.h"
#include "agpt.h"

class Material {
public:
	virtual ~Material() {}
	virtual agpt_material Export() const = 0;
};

class DisneyMaterial : public Material {
public:
	DisneyMaterial(const float3& color, float roughness, float metallic) {
		memset(&rec, 0, sizeof(rec));
		rec.type = AGPT_MAT_DISNEY;
		rec.eta = 1.5f;
		rec.roughness = roughness;
		rec.metallic = metallic;
		const float strans = 0.f;   // no transmission upstream (material.h:19)
		float diffuseWeight = (1 - metallic) * (1 - strans);
		if (diffuseWeight > 0) {    // diffuse + retro lobes exist iff metallic < 1 (material.h:27-36)
			float3 R = diffuseWeight * color;
			rec.diffuse_r[0] = R.x; rec.diffuse_r[1] = R.y; rec.diffuse_r[2] = R.z;
			rec.lobes |= AGPT_LOBE_DIFFUSE | AGPT_LOBE_RETRO;
		}
		const float aspect = 1.f;
		float ax = std::max(.001f, sqr(roughness) / aspect);
		float ay = std::max(.001f, sqr(roughness) * aspect);
		rec.alpha_x = std::max(0.001f, ax);   // clamped again by TrowbridgeReitzDistribution (microfacet.h:120-122)
		rec.alpha_y = std::max(0.001f, ay);
		const float specTint = 0.f;
		float r0 = sqr(rec.eta - 1) / sqr(rec.eta + 1);   // SchlickR0FromEta (disney.h:22)
		float3 Cspec0 = Lerp(metallic, r0 * Lerp(specTint, float3(1.f), float3(1.f)), color);
		rec.spec_r0[0] = Cspec0.x; rec.spec_r0[1] = Cspec0.y; rec.spec_r0[2] = Cspec0.z;
		rec.lobes |= AGPT_LOBE_MICROFACET;
	}
	agpt_material Export() const override { return rec; }
	static std::shared_ptr<DisneyMaterial> Make(const float3& color, float roughness, float metallic) {
		return std::make_shared<DisneyMaterial>(color, roughness, metallic);
	}
private:
	agpt_material rec;
};

class MirrorMaterial : public Material {
public:
	MirrorMaterial(const float3& r) {
		memset(&rec, 0, sizeof(rec));
		rec.type = AGPT_MAT_MIRROR;
		rec.lobes = AGPT_LOBE_SPECULAR;
		rec.mirror_r[0] = r.x; rec.mirror_r[1] = r.y; rec.mirror_r[2] = r.z;
	}
	agpt_material Export() const override { return rec; }
	static std::shared_ptr<MirrorMaterial> Make(const float3& r) { return std::make_shared<MirrorMaterial>(r); }
private:
	agpt_material rec;
};

// EXTENSION (not a reference class): a rough dielectric interface as PBRT-v3's GlassMaterial builds
// it -- MicrofacetReflection(Kr, TrowbridgeReitz(alpha), FresnelDielectric(1, eta)), which the reference's own classes
// can express (reflection.h:38-78, microfacet.h:112-153,220-228), plus MicrofacetTransmission(Kt, ...), which they
// cannot.  alpha = max(.001, roughness^2) like DisneyMaterial (material.h:39-41).  transmit = false keeps the
// reflection lobe alone (the half that can be checked against the reference).  See include/agpt.h AGPT_LOBE_GLASS_*.
#define AGPT_HAS_GLASS 1
class GlassMaterial : public Material {
public:
	GlassMaterial(const float3& Kr, const float3& Kt, float roughness, float eta = 1.5f, bool transmit = true) {
		memset(&rec, 0, sizeof(rec));
		rec.type = AGPT_MAT_GLASS;
		rec.lobes = AGPT_LOBE_GLASS_REFLECT | (transmit ? AGPT_LOBE_GLASS_TRANSMIT : 0);
		rec.roughness = roughness;
		rec.eta = eta;
		rec.mirror_r[0] = Kr.x; rec.mirror_r[1] = Kr.y; rec.mirror_r[2] = Kr.z;
		rec.diffuse_r[0] = Kt.x; rec.diffuse_r[1] = Kt.y; rec.diffuse_r[2] = Kt.z;
		rec.alpha_x = rec.alpha_y = std::max(0.001f, std::max(.001f, sqr(roughness)));
	}
	agpt_material Export() const override { return rec; }
	static std::shared_ptr<GlassMaterial> Make(const float3& Kr, const float3& Kt, float roughness, float eta = 1.5f, bool transmit = true) {
		return std::make_shared<GlassMaterial>(Kr, Kt, roughness, eta, transmit);
	}
private:
	agpt_material rec;
};
