/* TEST INFRASTRUCTURE — CPU restatement (plain C99) of the reference's small-step XPBD tet path.
 * See xpbd_oracle.c for the reference file:line each function follows.
 * Parity status: PINNED BY EXECUTION — tests/test_oracle_ref.py checks this restatement bit for bit
 * against the unmodified reference (oracle/_ref/libxpbd_ref_strict.so) and tests/test_golden.py
 * against vectors generated from that reference (tests/golden/, oracle/gen_golden.py). */
#ifndef XPBD_ORACLE_H
#define XPBD_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xo_scene xo_scene;

/* Reference Settings POD, byte-identical (Settings.h:79-102). */
typedef struct xo_settings {
	float timeScale, substepsPerSecond;
	uint32_t volumePasses, _pad0;
	float gravity[2];
	float compliance, damping, pbdDamping, drag, poissonsRatio, wonkiness, leftRightSeparation;
	uint32_t flags;
	float areaAndTimeCorrectedPbdDamping, volumeAndTimeCorrectedPbdDamping;
	float amortizedAreaAndTimeCorrectedPbdDamping, amortizedVolumeAndTimeCorrectedPbdDamping;
	float timeCorrectedDrag;
	uint32_t _pad1;
	float lockedRightTransform[4];
	float lockedRightTransform3d[12]; /* 3 columns of vec3 padded to 16 B */
	uint32_t tickId, _pad2[3];
} xo_settings;

typedef struct xo_manipulator {
	float pos[3], manipPlaneNormal[3], pick0[3], pickDir[3], pickDirOld[3], pickDirTarget[3];
	int32_t picked;
	uint32_t pickedPointIdx;
} xo_manipulator;

/* MeshGen.cpp:156-244.  nodes: 3*(w+1)(h+1)(d+1) floats; idxStream: 30*w*h*d u32 ([4,v0..v3]*). */
void xo_generate_tet_block(uint32_t width, uint32_t height, uint32_t depth, float sx, float sy, float sz,
                           uint32_t pattern, float wonkiness, float* nodes, uint32_t* idxStream);

/* GeoLinear3d::Init, Geo.cpp:697-772 (tets only). */
xo_scene* xo_create(const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream, uint32_t idxCount,
                    float density, int autoResize);
void xo_destroy(xo_scene* s);
uint32_t xo_vert_count(const xo_scene* s);
uint32_t xo_tet_count(const xo_scene* s);

void xo_get_order(const xo_scene* s, uint32_t* order);
void xo_set_order(xo_scene* s, const uint32_t* order);
void xo_get_state(const xo_scene* s, double* X, double* V, float* w);
void xo_set_state(xo_scene* s, const double* X, const double* V, const float* w);
void xo_get_rest(const xo_scene* s, double* X0, double* O, uint8_t* flags);
void xo_set_elements(xo_scene* s, const float* Qi9, const float* QQ3, const float* QR3, const float* volume, const float* area);
void xo_set_rest(xo_scene* s, const double* X0);
void xo_get_elements(const xo_scene* s, uint32_t* idx4, float* Qi9, float* QQ3, float* QR3, float* volume, float* area);
void xo_get_origin(const xo_scene* s, float* o);

/* Geo3d::Substep, Geo.cpp:305-356 (+ ground / handles extensions). */
void xo_substep(xo_scene* s, const xo_settings* settings, const xo_manipulator* manip, float dt, uint32_t n);
/* The same substep split into phases (predict | sweep of tOrder[begin,end) | post); used by the tests that emulate a
 * partitioned multi-rank run on CPU. */
void xo_phase_predict(xo_scene* s, const xo_settings* settings, float dt);
void xo_phase_sweep(xo_scene* s, const xo_settings* settings, float dt, uint32_t begin, uint32_t end);
void xo_phase_elems(xo_scene* s, const xo_settings* settings, float dt, int kind, const uint32_t* elems, uint32_t count);
void xo_phase_post(xo_scene* s, const xo_settings* settings, const xo_manipulator* manip, float dt);
void xo_set_flags(xo_scene* s, const uint8_t* flags);
void xo_set_ground(xo_scene* s, int enabled, float y0, float friction);
/* Study mode (tools/fp32_state_study.py): round X, O, V to fp32 after every write.  NOT the reference's algorithm. */
void xo_set_state_precision(xo_scene* s, int f32);
void xo_set_handles(xo_scene* s, uint32_t count, const uint32_t* vertIdx, const float* targetXYZ);

/* Geo3d::Transform, Geo.cpp:358-364. m9 column-major. */
void xo_transform(xo_scene* s, const float* m9);
/* GeoLinear3d::CalculateVolume, Geo.cpp:827-832. */
float xo_volume(const xo_scene* s);
/* Statistics for the 1000-frame trajectory checks (not in the reference; fp64 accumulation). */
void xo_energy(const xo_scene* s, const xo_settings* settings, double* kinetic, double* gravitational, double* elasticDev,
               double* elasticVol);

double xo_time_substeps(xo_scene* s, const xo_settings* settings, float dt, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
