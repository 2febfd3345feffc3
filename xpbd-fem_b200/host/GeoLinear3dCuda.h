// Reference-side adapter: a `Geo` (Geo.h:15-35) whose Substep runs on the B200 through the C ABI.
//
// Compile this header INSIDE the reference tree (it includes the reference's Geo.h); it is not part of
// libxpbd_fem_b200.so.  `Sim` only talks to `Geo*` through virtuals (Demo.cpp:86-88, 105-113, 156-162), so
// constructing a GeoLinear3dCuda instead of a GeoLinear3d in Sim::AddBlock (Demo.cpp:140-144) is the whole
// integration.  Geo3d::Substep is `final` (Geo.h:56), hence the adapter derives from Geo, not Geo3d.
//
// The test harness builds this adapter against the unmodified reference headers and
// tests/test_gpu_adapter.py drives both Geo implementations through the same virtual calls.
#pragma once

#include <float.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "Geo.h"
#include "xpbd_fem_b200.h"

struct GeoLinear3dCuda : public Geo {
	xf_scene* scene = nullptr;
	std::vector<double> hostX, hostV; // lazily refreshed mirror for Pick/Render
	std::vector<float> hostW;
	std::vector<uint8_t> hostFlags;
	bool mirrorValid = false;

	// Same signature as GeoLinear3d::Init (Geo.h:147) minus the Allocator: device memory is owned by the library.
	bool Init(float density, const float* nodeData, uint32_t nodeDataCount, const uint32_t* idxData, uint32_t idxDataCount, bool autoResize,
	          int device = 0, int precision = XF_PRECISION_EXACT, const uint32_t* colorHint = nullptr) {
		xf_create_params p;
		xf_default_create_params(&p);
		p.device = device;
		p.density = density;
		p.autoResize = autoResize ? 1 : 0;
		p.precision = precision;
		p.colorHint = colorHint;
		p.colorHintCount = colorHint ? idxDataCount / 5 : 0;
		if (xf_create(&p, nodeData, nodeDataCount, idxData, idxDataCount, &scene) != XF_OK) {
			fprintf(stderr, "GeoLinear3dCuda::Init: %s\n", xf_last_error());
			return false;
		}
		hostFlags.resize(xf_vert_count(scene));
		xf_get_rest(scene, nullptr, nullptr, hostFlags.data());
		return true;
	}
	~GeoLinear3dCuda() { xf_destroy(scene); }

	void Substep(const Settings& settings, const Manipulator& manip, float dt) final {
		static_assert(sizeof(Settings) == sizeof(xf_settings), "Settings POD must stay byte-identical to xf_settings");
		xf_manipulator m;
		memcpy(m.pos, &manip.pos, 12); memcpy(m.manipPlaneNormal, &manip.manipPlaneNormal, 12); memcpy(m.pick0, &manip.pick0, 12);
		memcpy(m.pickDir, &manip.pickDir, 12); memcpy(m.pickDirOld, &manip.pickDirOld, 12); memcpy(m.pickDirTarget, &manip.pickDirTarget, 12);
		m.picked = manip.pickedGeo == this ? 1 : 0; // Geo.cpp:334
		m.pickedPointIdx = manip.pickedPointIdx;
		if (xf_substep(scene, reinterpret_cast<const xf_settings*>(&settings), &m, dt, 1) != XF_OK) {
			fprintf(stderr, "GeoLinear3dCuda::Substep: %s\n", xf_last_error());
		}
		mirrorValid = false;
	}
	void Transform(mat3 t) final {
		const float m9[9] = { t[0][0], t[0][1], t[0][2], t[1][0], t[1][1], t[1][2], t[2][0], t[2][1], t[2][2] };
		xf_transform(scene, m9);
		mirrorValid = false;
	}
	// Constrain/Damp are sub-phases of Substep in the reference; on the device they are fused into xf_substep.
	void Constrain(const Settings&, float) final {}
	void Damp(const Settings&, float) final {}
	float CalculateVolume() const final {
		float v = 0.0f;
		xf_volume(scene, &v);
		return v;
	}
	uint32_t VertCount() const final { return xf_vert_count(scene); }
	uint32_t ElementCount() const final { return xf_element_count(scene); }

	void RefreshMirror() {
		if (mirrorValid) { return; }
		const uint32_t n = xf_vert_count(scene);
		hostX.resize(3 * (size_t)n); hostV.resize(3 * (size_t)n); hostW.resize(n);
		xf_get_state(scene, hostX.data(), hostV.data(), hostW.data());
		mirrorValid = true;
	}
	// Geo3d::Pick (Geo.cpp:366-385) on the host mirror, same selection rule so that a drag grabs the vertex the reference would:
	// candidates are pickable vertices that still have mass (w != 0: a locked vertex cannot be grabbed) in front of the ray origin;
	// they are ranked by t * distance-to-ray (t = distance along the ray), ties go to the LATER vertex, and the outputs are
	// only written when some candidate wins.
	void Pick(vec3 rayOrigin, vec3 rayDir, vec3* outNearestPoint, uint32_t* outNearestPointIdx, float* outDistance) final {
		RefreshMirror();
		float bestRank = FLT_MAX;
		for (uint32_t i = 0; i < hostFlags.size(); i++) {
			if (!(hostFlags[i] & Geo::Pickable) || hostW[i] == 0.0f) { continue; }
			const vec3 p = vec3(dvec3(hostX[3 * i], hostX[3 * i + 1], hostX[3 * i + 2]));
			float dist = FLT_MAX, rank = FLT_MAX;
			const float t = dot(rayDir, p - rayOrigin);
			if (t >= 0.0f) {
				dist = distance(p, rayOrigin + t * rayDir);
				rank = t * dist;
			}
			if (bestRank >= rank) {
				bestRank = rank;
				*outDistance = dist;
				*outNearestPoint = p;
				*outNearestPointIdx = i;
			}
		}
	}
	void Render(const Settings&) final { RefreshMirror(); /* draw from hostX with the host's renderer */ }
};
