// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Headless C-ABI harness around the UNMODIFIED reference sources of jak-xyz/xpbd-fem.  It is
// compiled together with Geo.cpp Fem.cpp Connectivity.cpp MeshGen.cpp Allocator.cpp
// vectormath.cpp DebugGeo.cpp *where they lie* under /root/reference/XPBDFEM (see
// oracle/Makefile); no reference source is copied into this repository.  The resulting
// shared objects live in oracle/_ref/ (git-ignored) and are used by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs only.
//
// What it drives (reference file:line):
//   GenerateBlock / GenerateArmadillo          MeshGen.cpp:246-260, 54-80
//   GeoLinear3d::Init                          Geo.cpp:697-772
//   Geo3d::Substep                             Geo.cpp:305-356
//   GeoLinear3d::CalculateVolume               Geo.cpp:827-832
//   Geo3d::Transform                           Geo.cpp:358-364
//   public members X, O, V, w, flags, t, tOrder (Geo.h:61-66, 158-163)
//
// Two extensions that the reference does not have are mirrored here so
// the CUDA path can be checked against *something*: a ground plane and multiple drag
// handles.  They need a restated Substep (ref_substep_ext below) because Geo3d::Substep is
// monolithic; with both extensions disabled ref_substep_ext is checked bit-for-bit against
// Geo3d::Substep by tests/test_oracle_ref.py.  Parity for the extensions themselves is
// "unpinned" by the reference.
#include "Demo.h"
#include "Geo.h"
#include "MeshGen.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>

extern "C" double performance_now() { return 0.0; } // Demo.cpp:14 (only feeds the HUD averages)

namespace {

struct ManipPod {  // mirror of include/xpbd_fem_b200.h : xf_manipulator
	float pos[3], manipPlaneNormal[3], pick0[3], pickDir[3], pickDirOld[3], pickDirTarget[3];
	int32_t picked;            // 0 = nothing picked (pickedGeo == nullptr)
	uint32_t pickedPointIdx;
};

struct RefScene {
	void* arena = nullptr;
	size_t arenaBytes = 0;
	Allocator alloc;
	GeoLinear3d* geo = nullptr;
	const float* nodeData = nullptr;
	uint32_t nodeDataCount = 0;
	const uint32_t* idxData = nullptr;
	uint32_t idxDataCount = 0;
	// extensions (x1, x2)
	int groundOn = 0;
	float groundY = 0.0f;
	float groundFriction = 0.0f;
	uint32_t handleCount = 0;
	uint32_t handleIdx[64];
	float handleTarget[64][3];
};

vec3 V3(const float* p) { return vec3(p[0], p[1], p[2]); }

Manipulator ToManip(const ManipPod* m, Geo* geo) {
	Manipulator out;
	if (!m) {
		out.pos = out.manipPlaneNormal = out.pick0 = out.pickDir = out.pickDirOld = out.pickDirTarget = vec3(0.0f);
		out.pickedGeo = nullptr;
		return out;
	}
	out.pos = V3(m->pos);
	out.manipPlaneNormal = V3(m->manipPlaneNormal);
	out.pick0 = V3(m->pick0);
	out.pickDir = V3(m->pickDir);
	out.pickDirOld = V3(m->pickDirOld);
	out.pickDirTarget = V3(m->pickDirTarget);
	out.pickedGeo = m->picked ? geo : nullptr;
	out.pickedPointIdx = m->pickedPointIdx;
	return out;
}

RefScene* NewScene(size_t tetGuess, size_t vertGuess) {
	RefScene* s = new RefScene();
	// T4 costs ~311 B/tet of arena in GeoLinear3d::Init; the block generators add
	// 9 u32/hex + 30 u32/hex + 3 f32/vert.  Be generous.
	s->arenaBytes = (size_t)64 * 1024 * 1024 + tetGuess * 640 + vertGuess * 512;
	s->arena = malloc(s->arenaBytes);
	if (!s->arena) { delete s; return nullptr; }
	s->alloc.Initialize(s->arena, s->arenaBytes);
	return s;
}

void Finish(RefScene* s, float density, bool autoResize) {
	s->geo = s->alloc.New<GeoLinear3d>();
	s->geo->Init(&s->alloc, density, s->nodeData, s->nodeDataCount, s->idxData, s->idxDataCount, autoResize);
	s->geo->volume0 = s->geo->CalculateVolume();
}

}  // namespace

extern "C" {

// GenerateBlock(Element_T4, w, h, vec2(sx, sy), pattern, wonkiness) -> GeoLinear3d::Init
void* ref_create_block(uint32_t width, uint32_t height, float scaleX, float scaleY, uint32_t pattern, float wonkiness, float density) {
	size_t hexes = (size_t)width * height * height;
	RefScene* s = NewScene(hexes * 6, (size_t)(width + 1) * (height + 1) * (height + 1));
	if (!s) { return nullptr; }
	GenerateBlock(Element_T4, &s->alloc, width, height, vec2(scaleX, scaleY), pattern, wonkiness, &s->nodeData, &s->nodeDataCount, &s->idxData, &s->idxDataCount);
	Finish(s, density, false);
	return s;
}

// Demo::UpdateSettings' Armadillo path: autoResize = true, density = 2 (Demo.cpp:123, 336)
void* ref_create_armadillo(float density) {
	RefScene* s = NewScene(4096, 4096);
	if (!s) { return nullptr; }
	GenerateArmadillo(Element_T4, &s->nodeData, &s->nodeDataCount, &s->idxData, &s->idxDataCount);
	Finish(s, density, true);
	return s;
}

// Arbitrary mesh in the reference's own stream format ([4, v0, v1, v2, v3]*).
void* ref_create_mesh(const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream, uint32_t idxCount, float density, int autoResize) {
	RefScene* s = NewScene(idxCount / 5, nodeFloatCount / 3);
	if (!s) { return nullptr; }
	float* n = s->alloc.Alloc<float>(nodeFloatCount);
	uint32_t* i = s->alloc.Alloc<uint32_t>(idxCount);
	memcpy(n, nodeXYZ, sizeof(float) * nodeFloatCount);
	memcpy(i, idxStream, sizeof(uint32_t) * idxCount);
	s->nodeData = n; s->nodeDataCount = nodeFloatCount;
	s->idxData = i; s->idxDataCount = idxCount;
	Finish(s, density, autoResize != 0);
	return s;
}

void ref_destroy(void* h) {
	RefScene* s = (RefScene*)h;
	if (!s) { return; }
	free(s->arena);
	delete s;
}

uint32_t ref_vert_count(void* h) { return ((RefScene*)h)->geo->vertCount; }
uint32_t ref_tet_count(void* h) { return ((RefScene*)h)->geo->con.tetCount; }
uint32_t ref_node_float_count(void* h) { return ((RefScene*)h)->nodeDataCount; }
uint32_t ref_idx_count(void* h) { return ((RefScene*)h)->idxDataCount; }
size_t ref_arena_used(void* h) { RefScene* s = (RefScene*)h; return (size_t)(s->alloc.head - s->alloc.mem); }

// The mesh exactly as it was handed to GeoLinear3d::Init, so the CUDA library can be fed the same bytes.
void ref_get_mesh(void* h, float* nodeXYZ, uint32_t* idxStream) {
	RefScene* s = (RefScene*)h;
	memcpy(nodeXYZ, s->nodeData, sizeof(float) * s->nodeDataCount);
	memcpy(idxStream, s->idxData, sizeof(uint32_t) * s->idxDataCount);
}

void ref_get_order(void* h, uint32_t* order) {
	RefScene* s = (RefScene*)h;
	memcpy(order, s->geo->tOrder, sizeof(uint32_t) * s->geo->con.tetCount);
}
// Inject the CUDA schedule's equivalent serial order (Geo.h:161 is a public member).
void ref_set_order(void* h, const uint32_t* order) {
	RefScene* s = (RefScene*)h;
	memcpy(s->geo->tOrder, order, sizeof(uint32_t) * s->geo->con.tetCount);
}

// State access: packed xyz triples of doubles (the reference's dvec3 is 32 B with padding).
void ref_get_state(void* h, double* X, double* V, float* w) {
	GeoLinear3d* g = ((RefScene*)h)->geo;
	for (uint32_t i = 0; i < g->vertCount; i++) {
		if (X) { X[3 * i + 0] = g->X[i].x; X[3 * i + 1] = g->X[i].y; X[3 * i + 2] = g->X[i].z; }
		if (V) { V[3 * i + 0] = g->V[i].x; V[3 * i + 1] = g->V[i].y; V[3 * i + 2] = g->V[i].z; }
		if (w) { w[i] = g->w[i]; }
	}
}
void ref_get_rest(void* h, double* X0, double* O, uint8_t* flags) {
	GeoLinear3d* g = ((RefScene*)h)->geo;
	for (uint32_t i = 0; i < g->vertCount; i++) {
		if (X0) { X0[3 * i + 0] = g->X0[i].x; X0[3 * i + 1] = g->X0[i].y; X0[3 * i + 2] = g->X0[i].z; }
		if (O) { O[3 * i + 0] = g->O[i].x; O[3 * i + 1] = g->O[i].y; O[3 * i + 2] = g->O[i].z; }
		if (flags) { flags[i] = g->flags[i]; }
	}
}
void ref_set_state(void* h, const double* X, const double* V, const float* w) {
	GeoLinear3d* g = ((RefScene*)h)->geo;
	for (uint32_t i = 0; i < g->vertCount; i++) {
		if (X) { g->X[i] = dvec3(X[3 * i + 0], X[3 * i + 1], X[3 * i + 2]); }
		if (V) { g->V[i] = dvec3(V[3 * i + 0], V[3 * i + 1], V[3 * i + 2]); }
		if (w) { g->w[i] = w[i]; }
	}
}
void ref_get_origin(void* h, float* o) {
	GeoLinear3d* g = ((RefScene*)h)->geo;
	o[0] = g->origin.x; o[1] = g->origin.y; o[2] = g->origin.z;
}

// Per-element constants as InitFiniteElement left them (Geo.h:89-92, Fem.h:52-60).
void ref_get_elements(void* h, uint32_t* idx4, float* Qi9, float* QQ3, float* QR3, float* volume, float* surfaceArea) {
	GeoLinear3d* g = ((RefScene*)h)->geo;
	for (uint32_t e = 0; e < g->con.tetCount; e++) {
		const T4& t = g->t[e];
		for (int j = 0; j < 4; j++) { if (idx4) { idx4[4 * e + j] = t.i[j]; } }
		// Qi9 is column-major like the reference's mat3: Qi9[3*c + r] = Qi[c][r]
		for (int c = 0; c < 3; c++) { for (int r = 0; r < 3; r++) { if (Qi9) { Qi9[9 * e + 3 * c + r] = t.ep.Qi[c][r]; } } }
		for (int j = 0; j < 3; j++) { if (QQ3) { QQ3[3 * e + j] = t.ep.ic.QQ[j]; } if (QR3) { QR3[3 * e + j] = t.ep.ic.QR[j]; } }
		if (volume) { volume[e] = t.ep.volume; }
		if (surfaceArea) { surfaceArea[e] = t.ep.surfaceArea; }
	}
}

// Geo3d::Transform (Geo.cpp:358-364); m9 is column-major mat3.
void ref_transform(void* h, const float* m9) {
	GeoLinear3d* g = ((RefScene*)h)->geo;
	g->Transform(mat3(V3(m9 + 0), V3(m9 + 3), V3(m9 + 6)));
}

float ref_volume(void* h) { return ((RefScene*)h)->geo->CalculateVolume(); }

// Geo::Pick through the virtual interface (Geo.h:24, Geo.cpp:366-385).  out5 = {point.xyz, distance, found}; idx = the vertex.
// The reference writes its outputs only when a candidate wins, so they are pre-set to sentinels here.
static void PickThrough(Geo* geo, const float* rayOrigin3, const float* rayDir3, float* out5, uint32_t* outIdx) {
	vec3 point(-1.0f, -1.0f, -1.0f);
	uint32_t idx = 0xffffffffu;
	float dist = -1.0f;
	geo->Pick(V3(rayOrigin3), V3(rayDir3), &point, &idx, &dist);
	out5[0] = point.x; out5[1] = point.y; out5[2] = point.z; out5[3] = dist; out5[4] = idx != 0xffffffffu ? 1.0f : 0.0f;
	*outIdx = idx;
}
void ref_pick(void* h, const float* rayOrigin3, const float* rayDir3, float* out5, uint32_t* outIdx) {
	PickThrough(((RefScene*)h)->geo, rayOrigin3, rayDir3, out5, outIdx);
}

// n calls of the reference's own Geo3d::Substep.  `settings160` is the reference's Settings
// POD verbatim (160 bytes, Settings.h:79-102); tickId is advanced per substep the way
// Sim::Update does (Demo.cpp:81, 89).
void ref_substep(void* h, const void* settings160, const void* manipPod, float dt, uint32_t n) {
	RefScene* s = (RefScene*)h;
	Settings settings;
	static_assert(sizeof(Settings) == 160, "Settings POD layout changed");
	memcpy((void*)&settings, settings160, sizeof(Settings));
	Manipulator manip = ToManip((const ManipPod*)manipPod, s->geo);
	for (uint32_t k = 0; k < n; k++) {
		s->geo->Substep(settings, manip, dt);
		settings.tickId++;
	}
}

// ---- extensions x1 / x2 (not in the reference; semantics defined in DESIGN.md) ----
void ref_set_ground(void* h, int enabled, float y0, float friction) {
	RefScene* s = (RefScene*)h;
	s->groundOn = enabled; s->groundY = y0; s->groundFriction = friction;
}
void ref_set_handles(void* h, uint32_t count, const uint32_t* vertIdx, const float* targetXYZ) {
	RefScene* s = (RefScene*)h;
	s->handleCount = count > 64 ? 64 : count;
	for (uint32_t k = 0; k < s->handleCount; k++) {
		s->handleIdx[k] = vertIdx[k];
		for (int j = 0; j < 3; j++) { s->handleTarget[k][j] = targetXYZ[3 * k + j]; }
	}
}

// Substep restated around the reference's own Constrain()/Damp() so the two extensions can be
// spliced in where DESIGN.md puts them (ground: after Constrain, before the locks; handles:
// after the single-manipulator projection).  With no extension enabled this must equal
// Geo3d::Substep bit for bit (checked by the tests).
void ref_substep_ext(void* h, const void* settings160, const void* manipPod, float dt, uint32_t n) {
	RefScene* s = (RefScene*)h;
	GeoLinear3d* g = s->geo;
	Settings settings;
	memcpy((void*)&settings, settings160, sizeof(Settings));
	Manipulator manip = ToManip((const ManipPod*)manipPod, g);
	for (uint32_t k = 0; k < n; k++) {
		for (uint32_t i = 0; i < g->vertCount; i++) {
			g->V[i] += dvec3(settings.gravity.x * dt, settings.gravity.y * dt, 0.0f);
			g->V[i] *= 1.0f - settings.timeCorrectedDrag;
			g->O[i] = g->X[i];
			g->X[i] += g->V[i] * dt;
		}
		g->Constrain(settings, dt);
		if (s->groundOn) {
			double y0 = (double)s->groundY;
			double keep = (double)(1.0f - s->groundFriction);
			for (uint32_t i = 0; i < g->vertCount; i++) {
				if (g->X[i].y < y0) {
					g->X[i].y = y0;
					g->X[i].x = g->O[i].x + (g->X[i].x - g->O[i].x) * keep;
					g->X[i].z = g->O[i].z + (g->X[i].z - g->O[i].z) * keep;
				}
			}
		}
		if (settings.flags & Settings_LockLeft) {
			for (uint32_t i = 0; i < g->vertCount; i++) {
				if (g->flags[i] & Geo::Left) { g->X[i] = g->O[i]; g->w[i] = 0.0f; }
			}
		}
		if (settings.flags & Settings_LockRight) {
			for (uint32_t i = 0; i < g->vertCount; i++) {
				if (g->flags[i] & Geo::Right) {
					g->X[i] = g->O[i] = dvec3(g->origin + (settings.lockedRightTransform3d * vec3(g->X0[i])));
					g->w[i] = 0.0f;
				}
			}
		}
		if (manip.pickedGeo == g) {
			uint32_t i = manip.pickedPointIdx;
			float t = dot(manip.manipPlaneNormal, manip.pick0 - manip.pos) / dot(manip.manipPlaneNormal, manip.pickDirTarget);
			vec3 target = manip.pos + t * manip.pickDirTarget;
			g->X[i] += dvec3((target - vec3(g->X[i])) * (g->w[i] / (max(0.000001f, g->w[i]) + 1.8f / (dt * dt))));
		}
		for (uint32_t hd = 0; hd < s->handleCount; hd++) {
			uint32_t i = s->handleIdx[hd];
			vec3 target = V3(s->handleTarget[hd]);
			g->X[i] += dvec3((target - vec3(g->X[i])) * (g->w[i] / (max(0.000001f, g->w[i]) + 1.8f / (dt * dt))));
		}
		for (uint32_t i = 0; i < g->vertCount; i++) {
			g->V[i] = (g->X[i] - g->O[i]) * (1.0f / dt);
		}
		uint32_t rayleighDampingType = (settings.flags >> Settings_RayleighTypeBit) & Settings_RayleighTypeMask;
		if (rayleighDampingType == Rayleigh_PostAmortized) {
			Settings amortizedSettings = settings;
			amortizedSettings.damping *= (float)AmortizationPeriod;
			amortizedSettings.volumeAndTimeCorrectedPbdDamping = settings.amortizedVolumeAndTimeCorrectedPbdDamping;
			g->Damp(amortizedSettings, dt);
		} else {
			g->Damp(settings, dt);
		}
		settings.tickId++;
	}
}

#ifdef XF_WITH_CUDA_ADAPTER
}  // extern "C"
// ---- drop-in demonstration: the product's reference-side adapter driven through the reference's own `Geo`
// virtual interface, exactly like Sim::Update does (Demo.cpp:86-88). Built only into libxpbd_ref_adapter.so. ----
#include "GeoLinear3dCuda.h"
extern "C" {
struct AdapterScene {
	Geo* geo = nullptr;            // virtual calls only
	GeoLinear3dCuda* impl = nullptr;
};
void* ref_adapter_create(const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream, uint32_t idxCount, float density, int autoResize,
                         const uint32_t* colorHint) {
	AdapterScene* a = new AdapterScene();
	a->impl = new GeoLinear3dCuda();
	if (!a->impl->Init(density, nodeXYZ, nodeFloatCount, idxStream, idxCount, autoResize != 0, 0, XF_PRECISION_EXACT, colorHint)) {
		delete a->impl; delete a; return nullptr;
	}
	a->geo = a->impl;
	a->geo->volume0 = a->geo->CalculateVolume();
	return a;
}
void ref_adapter_destroy(void* h) { AdapterScene* a = (AdapterScene*)h; if (a) { delete a->impl; delete a; } }
uint32_t ref_adapter_vert_count(void* h) { return ((AdapterScene*)h)->geo->VertCount(); }
uint32_t ref_adapter_element_count(void* h) { return ((AdapterScene*)h)->geo->ElementCount(); }
void ref_adapter_get_order(void* h, uint32_t* order) { xf_get_order(((AdapterScene*)h)->impl->scene, order); }
void ref_adapter_substep(void* h, const void* settings160, const void* manipPod, float dt, uint32_t n) {
	AdapterScene* a = (AdapterScene*)h;
	Settings settings;
	memcpy((void*)&settings, settings160, sizeof(Settings));
	Manipulator manip = ToManip((const ManipPod*)manipPod, a->geo);
	for (uint32_t k = 0; k < n; k++) {
		a->geo->Substep(settings, manip, dt); // virtual dispatch, one substep per call like the reference
		settings.tickId++;
	}
}
float ref_adapter_volume(void* h) { return ((AdapterScene*)h)->geo->CalculateVolume(); }
void ref_adapter_pick(void* h, const float* rayOrigin3, const float* rayDir3, float* out5, uint32_t* outIdx) {
	PickThrough(((AdapterScene*)h)->geo, rayOrigin3, rayDir3, out5, outIdx);
}
void ref_adapter_transform(void* h, const float* m9) { ((AdapterScene*)h)->geo->Transform(mat3(V3(m9 + 0), V3(m9 + 3), V3(m9 + 6))); }
void ref_adapter_get_state(void* h, double* X, double* V, float* w) {
	AdapterScene* a = (AdapterScene*)h;
	a->impl->RefreshMirror();
	const size_t n = a->impl->hostW.size();
	if (X) { memcpy(X, a->impl->hostX.data(), sizeof(double) * 3 * n); }
	if (V) { memcpy(V, a->impl->hostV.data(), sizeof(double) * 3 * n); }
	if (w) { memcpy(w, a->impl->hostW.data(), sizeof(float) * n); }
}
#endif

// ---- the reference's own frame driver: Sim (Demo.h:18-52) with one T4 block, stepped by Sim::Update ----
struct RefSim {
	Sim* sim = nullptr;
	Manipulator manip;
};
void* ref_sim_create(const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream, uint32_t idxCount, const void* settings160, int autoResize) {
	RefSim* r = new RefSim();
	r->sim = new Sim();
	r->sim->Reset();
	memcpy((void*)&r->sim->settings, settings160, sizeof(Settings));
	r->sim->AddBlock(Element_T4, nodeXYZ, nodeFloatCount, idxStream, idxCount, autoResize != 0);
	r->sim->FinishAddingBlocks(); // Transform (rotate / offset) + volume0, Demo.cpp:156-168
	r->manip = ToManip(nullptr, nullptr);
	return r;
}
void ref_sim_destroy(void* h) { RefSim* r = (RefSim*)h; if (r) { delete r->sim; delete r; } }
void ref_sim_set_order(void* h, const uint32_t* order) {
	GeoLinear3d* g = (GeoLinear3d*)((RefSim*)h)->sim->geos[0];
	memcpy(g->tOrder, order, sizeof(uint32_t) * g->con.tetCount);
}
void ref_sim_get_state(void* h, double* X, double* V, float* w) {
	GeoLinear3d* g = (GeoLinear3d*)((RefSim*)h)->sim->geos[0];
	for (uint32_t i = 0; i < g->vertCount; i++) {
		if (X) { X[3 * i + 0] = g->X[i].x; X[3 * i + 1] = g->X[i].y; X[3 * i + 2] = g->X[i].z; }
		if (V) { V[3 * i + 0] = g->V[i].x; V[3 * i + 1] = g->V[i].y; V[3 * i + 2] = g->V[i].z; }
		if (w) { w[i] = g->w[i]; }
	}
}
// One Sim::Update.  settings160 / manipPod are in-out like the reference's members; returns the tick advance (= substeps).
uint32_t ref_sim_update(void* h, void* settings160, void* manipPod, float dt, float medianFrameTime) {
	RefSim* r = (RefSim*)h;
	Sim* sim = r->sim;
	// the UI writes sim.settings every frame (Demo::UpdateSettings, Demo.cpp:342); keep Sim's auto-updated members
	Settings in;
	memcpy((void*)&in, settings160, sizeof(Settings));
	sim->settings = in;
	ManipPod* mp = (ManipPod*)manipPod;
	Manipulator manip = ToManip(mp, sim->geos[0]);
	const uint32_t tick0 = sim->tickId;
	sim->Update(dt, medianFrameTime, &manip);
	memcpy(settings160, (void*)&sim->settings, sizeof(Settings));
	if (mp) { mp->pickDirTarget[0] = manip.pickDirTarget.x; mp->pickDirTarget[1] = manip.pickDirTarget.y; mp->pickDirTarget[2] = manip.pickDirTarget.z; }
	return sim->tickId - tick0;
}

// ---- Sim with several geos, and the block Demo::UpdateSettings builds from a Settings block (shape table, Demo.cpp:289-318) ----
struct RefMulti {
	Demo* demo = nullptr; // owns the Sim when the block came from Demo::UpdateSettings
	Sim* sim = nullptr;
};
void* ref_multi_create(const void* settings160) {
	RefMulti* r = new RefMulti();
	r->sim = new Sim();
	r->sim->Reset();
	memcpy((void*)&r->sim->settings, settings160, sizeof(Settings));
	return r;
}
// Demo::UpdateSettings(0, settings) on a fresh Demo: generates the block the settings describe, AddBlock, FinishAddingBlocks
void* ref_multi_from_settings(const void* settings160) {
	RefMulti* r = new RefMulti();
	r->demo = new Demo();
	r->demo->sims[0].Reset();
	r->demo->sims[1].Reset();
	Settings st;
	memcpy((void*)&st, settings160, sizeof(Settings));
	st.flags |= Settings_ConstructBlockFromSettings | Settings_ForceResetBlock;
	r->demo->UpdateSettings(0, st);
	r->sim = &r->demo->sims[0];
	return r;
}
void ref_multi_destroy(void* h) {
	RefMulti* r = (RefMulti*)h;
	if (!r) { return; }
	if (r->demo) { delete r->demo; } else { delete r->sim; }
	delete r;
}
void ref_multi_add_block(void* h, const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream, uint32_t idxCount, int autoResize) {
	((RefMulti*)h)->sim->AddBlock(Element_T4, nodeXYZ, nodeFloatCount, idxStream, idxCount, autoResize != 0);
}
void ref_multi_finish(void* h) { ((RefMulti*)h)->sim->FinishAddingBlocks(); }
void ref_multi_set_geo_offset(void* h, float x, float y) { ((RefMulti*)h)->sim->SetGeoOffset(vec2(x, y)); }
uint32_t ref_multi_geo_count(void* h) { return ((RefMulti*)h)->sim->geoCount; }
void ref_multi_geo_sizes(void* h, uint32_t geo, uint32_t* nV, uint32_t* nT) {
	GeoLinear3d* g = (GeoLinear3d*)((RefMulti*)h)->sim->geos[geo];
	*nV = g->vertCount;
	*nT = g->con.tetCount;
}
float ref_multi_volume0(void* h, uint32_t geo) { return ((RefMulti*)h)->sim->geos[geo]->volume0; }
void ref_multi_set_order(void* h, uint32_t geo, const uint32_t* order) {
	GeoLinear3d* g = (GeoLinear3d*)((RefMulti*)h)->sim->geos[geo];
	memcpy(g->tOrder, order, sizeof(uint32_t) * g->con.tetCount);
}
void ref_multi_get_mesh(void* h, uint32_t geo, double* X0, uint32_t* idx4) {
	GeoLinear3d* g = (GeoLinear3d*)((RefMulti*)h)->sim->geos[geo];
	for (uint32_t i = 0; i < g->vertCount; i++) { X0[3 * i + 0] = g->X0[i].x; X0[3 * i + 1] = g->X0[i].y; X0[3 * i + 2] = g->X0[i].z; }
	for (uint32_t t = 0; t < g->con.tetCount; t++) { for (int j = 0; j < 4; j++) { idx4[4 * t + j] = g->t[t].i[j]; } }
}
void ref_multi_get_state(void* h, uint32_t geo, double* X, double* V, float* w) {
	GeoLinear3d* g = (GeoLinear3d*)((RefMulti*)h)->sim->geos[geo];
	for (uint32_t i = 0; i < g->vertCount; i++) {
		if (X) { X[3 * i + 0] = g->X[i].x; X[3 * i + 1] = g->X[i].y; X[3 * i + 2] = g->X[i].z; }
		if (V) { V[3 * i + 0] = g->V[i].x; V[3 * i + 1] = g->V[i].y; V[3 * i + 2] = g->V[i].z; }
		if (w) { w[i] = g->w[i]; }
	}
}
// One Sim::Update; the manipulator holds geo `pickedGeo` (-1: nothing)
uint32_t ref_multi_update(void* h, void* settings160, void* manipPod, int pickedGeo, float dt, float medianFrameTime) {
	Sim* sim = ((RefMulti*)h)->sim;
	Settings in;
	memcpy((void*)&in, settings160, sizeof(Settings));
	sim->settings = in;
	ManipPod* mp = (ManipPod*)manipPod;
	Manipulator manip = ToManip(mp, pickedGeo >= 0 ? sim->geos[pickedGeo] : nullptr);
	const uint32_t tick0 = sim->tickId;
	sim->Update(dt, medianFrameTime, &manip);
	memcpy(settings160, (void*)&sim->settings, sizeof(Settings));
	if (mp) { mp->pickDirTarget[0] = manip.pickDirTarget.x; mp->pickDirTarget[1] = manip.pickDirTarget.y; mp->pickDirTarget[2] = manip.pickDirTarget.z; }
	return sim->tickId - tick0;
}

// Timing leg for bench.py: run `n` reference substeps and return elapsed seconds.
double ref_time_substeps(void* h, const void* settings160, float dt, uint32_t n) {
	auto t0 = std::chrono::steady_clock::now();
	ref_substep(h, settings160, nullptr, dt, n);
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
