// The reference's Sim (Demo.h:18-52) behind the C ABI: a container of Geos that share one Settings block and one frame clock.
//   Sim::Reset / AddBlock / FinishAddingBlocks / SetGeoOffset / Update      Demo.cpp:37-186
//   Demo::UpdateSettings' block table (shape -> hexes per side, size)        Demo.cpp:289-318
// Host-only logic (compiled with -ffp-contract=off: the offsets and scales must round like the reference's); every Geo is an
// ordinary xf_scene on the creating parameters' device and stream, stepped through xf_substep / xf_substep_varying.
#include <cstring>
#include <vector>

#include "xf_scene.h"

struct xf_sim {
	xf_create_params params;
	std::vector<xf_scene*> geos;
	std::vector<float> volume0;
	xf_frame_state state;
	float geoOffset[2] = { 0.0f, 0.0f };
};

namespace {
const uint32_t kMaxGeos = 32; // Sim::kMaxGeos, Demo.h:33
const float kSpacing = (20.0f / 100.0f) / (float)(31); // Demo.cpp:16

// transform(mat2 rot, vec2 offset) as the column-major mat3 Geo::Transform takes
void Affine(const float rot[4], float ox, float oy, float m9[9]) {
	m9[0] = rot[0]; m9[1] = rot[1]; m9[2] = 0.0f;
	m9[3] = rot[2]; m9[4] = rot[3]; m9[5] = 0.0f;
	m9[6] = ox; m9[7] = oy; m9[8] = 1.0f;
}
}  // namespace

extern "C" {

int xf_sim_create(const xf_create_params* params, xf_sim** outSim) {
	if (!params || !outSim) { return xf::Fail(XF_ERR_INVALID, "null argument"); }
	xf_sim* sim = new xf_sim();
	sim->params = *params;
	sim->params.colorHint = nullptr; // a hint belongs to one mesh; blocks built from settings bring their own
	sim->params.colorHintCount = 0;
	xf_frame_state_init(&sim->state);
	*outSim = sim;
	return XF_OK;
}

// Sim::Reset, Demo.cpp:170-177
int xf_sim_reset(xf_sim* sim) {
	if (!sim) { return xf::Fail(XF_ERR_INVALID, "null sim"); }
	for (xf_scene* g : sim->geos) { xf_destroy(g); }
	sim->geos.clear();
	sim->volume0.clear();
	xf_frame_state_init(&sim->state);
	return XF_OK;
}

int xf_sim_destroy(xf_sim* sim) {
	if (!sim) { return XF_OK; }
	xf_sim_reset(sim);
	delete sim;
	return XF_OK;
}

// Sim::AddBlock, Demo.cpp:120-154 (tet blocks only; the other element types are outside this library's path)
int xf_sim_add_block(xf_sim* sim, uint32_t elementType, const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream,
                     uint32_t idxCount, int autoResize, const uint32_t* colorHint, uint32_t colorHintCount) {
	if (!sim) { return xf::Fail(XF_ERR_INVALID, "null sim"); }
	if (elementType != XF_ELEMENT_T4) { return xf::Fail(XF_ERR_UNSUPPORTED, "only Element_T4 blocks are on the path this library implements"); }
	if (sim->geos.size() >= kMaxGeos) { return xf::Fail(XF_ERR_INVALID, "a Sim holds at most 32 geos (Demo.h:33)"); }
	xf_create_params p = sim->params;
	p.density = autoResize ? 2.0f : 1.0f; // Demo.cpp:123
	p.autoResize = autoResize ? 1 : 0;
	p.colorHint = colorHint;
	p.colorHintCount = colorHintCount;
	xf_scene* g = nullptr;
	int rc = xf_create(&p, nodeXYZ, nodeFloatCount, idxStream, idxCount, &g);
	if (rc != XF_OK) { return rc; }
	sim->geos.push_back(g);
	sim->volume0.push_back(0.0f);
	return XF_OK;
}

// Hexes per side and size of the block a Settings block asks for: the table of Demo::UpdateSettings, Demo.cpp:289-318, for
// 3D linear elements (width and height halved, size doubled, Demo.cpp:311-314).
int xf_block_from_settings(const xf_settings* st, uint32_t* outWidth, uint32_t* outHeight, float* outScaleX, float* outScaleY, uint32_t* outPattern) {
	if (!st) { return xf::Fail(XF_ERR_INVALID, "null settings"); }
	const uint32_t type = st->flags & 15u, shape = (st->flags >> 12) & 15u;
	if (type != XF_ELEMENT_T4) { return xf::Fail(XF_ERR_UNSUPPORTED, "only Element_T4 blocks are on the path this library implements"); }
	if (shape == 12u) { return xf::Fail(XF_ERR_UNSUPPORTED, "Shape_Armadillo is the caller's mesh asset: pass it to xf_sim_add_block with autoResize = 1"); }
	uint32_t width, height;
	float dimX, dimY;
	switch (shape) {
	case 0: width = 1; height = 1; dimX = dimY = 2.0f; break;           // Shape_Single
	case 1: width = 12; height = 1; dimX = dimY = 2.0f; break;          // Shape_Line
	case 2: width = 16; height = 4; dimX = dimY = 1.5f; break;          // Shape_BeamL
	case 3: width = 32; height = 8; dimX = dimY = 0.75f; break;         // Shape_BeamM
	case 4: width = 48; height = 12; dimX = dimY = 0.5f; break;         // Shape_BeamH
	case 5: width = 16; height = 8; dimX = 1.5f; dimY = 0.75f; break;   // Shape_BeamL1x2
	case 6: width = 32; height = 4; dimX = 0.75f; dimY = 1.5f; break;   // Shape_BeamL2x1
	case 7: width = 64; height = 4; dimX = 0.375f; dimY = 1.5f; break;  // Shape_BeamL4x1
	case 8: width = 128; height = 4; dimX = 0.1875f; dimY = 1.5f; break; // Shape_BeamL8x1
	case 9: width = height = 16; dimX = dimY = 1.5f; break;             // Shape_BoxL
	case 10: width = height = 24; dimX = dimY = 1.0f; break;            // Shape_BoxM
	default: width = height = 48; dimX = dimY = 0.5f; break;            // Shape_BoxH
	}
	width /= 2; height /= 2; // is3D
	dimX *= 2.0f; dimY *= 2.0f;
	if (width < 1) { width = 1; }
	if (height < 1) { height = 1; }
	if (outWidth) { *outWidth = width; }
	if (outHeight) { *outHeight = height; }
	if (outScaleX) { *outScaleX = (0.7f * kSpacing) * dimX; }
	if (outScaleY) { *outScaleY = (0.7f * kSpacing) * dimY; }
	if (outPattern) { *outPattern = (st->flags >> 16) & 3u; }
	return XF_OK;
}

// The block Demo::UpdateSettings builds for these settings (GenerateBlock -> GenerateTetBlock(width, height, height, scale.xyy),
// MeshGen.cpp:246-262), added like Sim::AddBlock does.
int xf_sim_add_block_from_settings(xf_sim* sim, const xf_settings* st) {
	if (!sim) { return xf::Fail(XF_ERR_INVALID, "null sim"); }
	uint32_t w = 0, h = 0, pattern = 0;
	float sx = 0.0f, sy = 0.0f;
	int rc = xf_block_from_settings(st, &w, &h, &sx, &sy, &pattern);
	if (rc != XF_OK) { return rc; }
	const size_t hexes = (size_t)w * h * h;
	std::vector<float> nodes(3 * (size_t)(w + 1) * (h + 1) * (h + 1));
	std::vector<uint32_t> idx(30 * hexes), hint(6 * hexes);
	rc = xf_generate_tet_block(w, h, h, sx, sy, sy, pattern, st->wonkiness, nodes.data(), idx.data(), hint.data());
	if (rc != XF_OK) { return xf::Fail(rc, "xf_generate_tet_block failed"); }
	const bool lattice = pattern == 0; // the 24-class hint is valid for Pattern_Uniform only
	return xf_sim_add_block(sim, XF_ELEMENT_T4, nodes.data(), (uint32_t)nodes.size(), idx.data(), (uint32_t)idx.size(), 0, lattice ? hint.data() : nullptr,
	                        lattice ? (uint32_t)hint.size() : 0u);
}

// Sim::FinishAddingBlocks, Demo.cpp:156-168: spread the geos vertically, rotate if asked, remember the rest volumes
int xf_sim_finish_adding_blocks(xf_sim* sim, const xf_settings* st) {
	if (!sim || !st) { return xf::Fail(XF_ERR_INVALID, "null argument"); }
	const uint32_t n = (uint32_t)sim->geos.size();
	for (uint32_t i = 0; i < n; i++) {
		const float yOff = (float)i * -0.03f + (float)(n - 1) * 0.015f;
		const bool rotate = (st->flags & (1u << 24)) != 0; // Settings_Rotate90Degrees
		float rot[4] = { 1.0f, 0.0f, 0.0f, 1.0f };        // columns
		if (rotate) { rot[0] = 0.0f; rot[1] = -1.0f; rot[2] = 1.0f; rot[3] = 0.0f; }
		// rot * vec2(0, yOff): row dots, vectormath.h
		const float rx = rot[0] * 0.0f + rot[2] * yOff, ry = rot[1] * 0.0f + rot[3] * yOff;
		float m9[9];
		Affine(rot, sim->geoOffset[0] + rx, sim->geoOffset[1] + ry, m9);
		if (sim->params.device >= 0) {
			int rc = xf_transform(sim->geos[i], m9);
			if (rc != XF_OK) { return rc; }
			rc = xf_volume(sim->geos[i], &sim->volume0[i]);
			if (rc != XF_OK) { return rc; }
		}
	}
	return XF_OK;
}

// Sim::SetGeoOffset, Demo.cpp:179-186
int xf_sim_set_geo_offset(xf_sim* sim, float x, float y) {
	if (!sim) { return xf::Fail(XF_ERR_INVALID, "null sim"); }
	if (sim->geoOffset[0] == x && sim->geoOffset[1] == y) { return XF_OK; }
	const float rot[4] = { 1.0f, 0.0f, 0.0f, 1.0f };
	float m9[9];
	Affine(rot, x - sim->geoOffset[0], y - sim->geoOffset[1], m9);
	for (xf_scene* g : sim->geos) {
		int rc = xf_transform(g, m9);
		if (rc != XF_OK) { return rc; }
	}
	sim->geoOffset[0] = x;
	sim->geoOffset[1] = y;
	return XF_OK;
}

// Sim::Update, Demo.cpp:37-103: one frame for every geo; `pickedGeo` = index of the geo the manipulator holds (-1: none)
int xf_sim_update(xf_sim* sim, xf_settings* settings, xf_manipulator* manip, int pickedGeo, float dt, float medianFrameTime, uint32_t* outSubsteps) {
	if (!sim) { return xf::Fail(XF_ERR_INVALID, "null sim"); }
	return xf::FrameUpdateGeos(sim->geos.data(), (uint32_t)sim->geos.size(), pickedGeo, settings, manip, dt, medianFrameTime, &sim->state, outSubsteps);
}

uint32_t xf_sim_geo_count(const xf_sim* sim) { return sim ? (uint32_t)sim->geos.size() : 0u; }
xf_scene* xf_sim_geo(xf_sim* sim, uint32_t i) { return (sim && i < sim->geos.size()) ? sim->geos[i] : nullptr; }
float xf_sim_volume0(const xf_sim* sim, uint32_t i) { return (sim && i < sim->volume0.size()) ? sim->volume0[i] : 0.0f; }
int xf_sim_get_frame_state(const xf_sim* sim, xf_frame_state* out) {
	if (!sim || !out) { return xf::Fail(XF_ERR_INVALID, "null argument"); }
	*out = sim->state;
	return XF_OK;
}

}  // extern "C"
