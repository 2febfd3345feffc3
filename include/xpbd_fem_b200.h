/*
 * xpbd_fem_b200.h — C ABI of the B200-native small-step XPBD tet solver.
 *
 * Drop-in boundary for the reference's `Geo` interface on the linear-tet path
 * (reference = jak-xyz/xpbd-fem, paths below relative to XPBDFEM/):
 *
 *   reference                                              this library
 *   ------------------------------------------------------ ------------------------------
 *   GenerateBlock(Element_T4, ...)        MeshGen.cpp:246   xf_generate_tet_block
 *   GeoLinear3d::Init                     Geo.cpp:697       xf_create
 *   Geo::Substep (Geo3d::Substep)         Geo.h:21, Geo.cpp:305   xf_substep
 *   Geo::Transform                        Geo.h:22, Geo.cpp:358   xf_transform
 *   Geo::CalculateVolume                  Geo.h:28, Geo.cpp:827   xf_volume
 *   Geo::VertCount / ElementCount         Geo.h:29-30       xf_vert_count / xf_element_count
 *   public members X, V, w, X0, O, flags  Geo.h:61-66       xf_get_state / xf_set_state / xf_get_rest
 *   public member tOrder                  Geo.h:161         xf_get_order (the schedule's equivalent serial order)
 *   Settings POD                          Settings.h:79-102 xf_settings (byte-identical, 160 B)
 *   Manipulator POD                       Manipulator.h:9-13 xf_manipulator (flat floats)
 *   wasm C exports (precedent for a flat C ABI)  wasm/main.cpp:8-17
 *
 * All entry points are plain C: pointers, sizes, PODs.  No C++ or torch types cross the boundary.
 * Every function returns 0 on success or a negative xf_status; xf_last_error() gives the message of
 * the calling thread's last failure.  The library never falls back to a CPU path: without a CUDA
 * device every compute entry point fails with XF_ERR_CUDA.
 *
 * Threading: one host thread per scene at a time.  Device work is enqueued on the scene's stream
 * (library-owned unless xf_create_params.stream is given); xf_substep is asynchronous, the state
 * getters synchronise.
 */
#ifndef XPBD_FEM_B200_H
#define XPBD_FEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: xf_info grew (chainedPermille), XF_GROUPING_CHAINS, xf_get_chain_info */
#define XF_ABI_VERSION 2

typedef enum xf_status {
	XF_OK = 0,
	XF_ERR_INVALID = -1,     /* bad argument / malformed mesh stream */
	XF_ERR_CUDA = -2,        /* CUDA runtime failure (message has the CUDA error string) */
	XF_ERR_UNSUPPORTED = -3, /* settings select something outside the tet hot path */
	XF_ERR_NOMEM = -4,
	XF_ERR_COLORING = -5     /* supplied colouring is not conflict-free, or too many colours */
} xf_status;

/* ---- flag word, identical to Settings.h:9-75 ---- */
#define XF_SETTINGS_ENERGY_BIT 6
#define XF_SETTINGS_ENERGY_MASK 31u
#define XF_SETTINGS_XPBD_SOLVE_BIT 11
#define XF_SETTINGS_RAYLEIGH_TYPE_BIT 20
#define XF_SETTINGS_RAYLEIGH_TYPE_MASK 3u
#define XF_SETTINGS_LOCK_LEFT (1u << 26)
#define XF_SETTINGS_LOCK_RIGHT (1u << 27)
#define XF_ELEMENT_T4 5u
#define XF_ENERGY_MIXED 3u
#define XF_ENERGY_MIXED_SEL 4u
#define XF_ENERGY_YEOH_SKIN 5u
#define XF_ENERGY_YEOH_SKIN_FAST 7u
#define XF_PATTERN_UNIFORM 0u
#define XF_PATTERN_MIRRORED 1u
#define XF_RAYLEIGH_PAPER 0u
#define XF_RAYLEIGH_LIMIT 1u
#define XF_RAYLEIGH_POST 2u
#define XF_RAYLEIGH_POST_AMORTIZED 3u
#define XF_AMORTIZATION_PERIOD 8u
/* per-vertex flags, Geo.h:16-20 */
#define XF_VERT_LEFT 1u
#define XF_VERT_RIGHT 2u
#define XF_VERT_PICKABLE 16u

/* Byte-identical to the reference's `struct Settings` (Settings.h:79-102; sizeof == 160, align 16). */
typedef struct xf_settings {
	float timeScale;
	float substepsPerSecond;
	uint32_t volumePasses;
	uint32_t _pad0;
	float gravity[2];
	float compliance;
	float damping;
	float pbdDamping;
	float drag;
	float poissonsRatio;
	float wonkiness;
	float leftRightSeparation;
	uint32_t flags;
	float areaAndTimeCorrectedPbdDamping;
	float volumeAndTimeCorrectedPbdDamping;
	float amortizedAreaAndTimeCorrectedPbdDamping;
	float amortizedVolumeAndTimeCorrectedPbdDamping;
	float timeCorrectedDrag;
	uint32_t _pad1;
	float lockedRightTransform[4];    /* mat2, column-major */
	float lockedRightTransform3d[12]; /* mat3, column-major, each column padded to 16 B */
	uint32_t tickId;
	uint32_t _pad2[3];
} xf_settings;

/* How xf_part_create splits the elements of one mesh over the ranks. */
typedef enum xf_partition_method {
	XF_PARTITION_SLABS = 0, /* contiguous chunks of the order "rest-pose centroid x": slabs for MeshGen blocks; a vertex has at most
	                           two copies, so the barrier-free cross-GPU schedule (k_part_dataflow) applies */
	XF_PARTITION_GRAPH = 1  /* greedy graph growing over the element adjacency graph (breadth-first from peripheral seeds, balanced
	                           part sizes): connected parts for irregular meshes; vertices may have copies on several ranks, which
	                           runs on the flag protocol (phases ordered by system-scope flags) */
} xf_partition_method;

/* Flat mirror of Manipulator.h:9-13.  picked != 0 <=> `manip.pickedGeo == this` (Geo.cpp:334). */
typedef struct xf_manipulator {
	float pos[3], manipPlaneNormal[3], pick0[3], pickDir[3], pickDirOld[3], pickDirTarget[3];
	int32_t picked;
	uint32_t pickedPointIdx;
} xf_manipulator;

typedef enum xf_precision {
	XF_PRECISION_EXACT = 0, /* every fp32 operation rounded separately, IEEE division: bit-identical to the
	                           reference built with -ffp-contract=off */
	XF_PRECISION_FAST = 1   /* same formulas, FMA contraction and prefactored constants recomputed in registers */
} xf_precision;

typedef enum xf_schedule {
	XF_SCHEDULE_AUTO = 0,
	XF_SCHEDULE_LAUNCH_PER_COLOR = 1, /* one kernel launch per colour phase */
	XF_SCHEDULE_PERSISTENT = 2,       /* one cooperative launch per xf_substep call, grid barriers between colours */
	XF_SCHEDULE_BRICKS = 3,           /* retired in round 2 (measured slower than PERSISTENT): accepted, runs as XF_SCHEDULE_PERSISTENT */
	XF_SCHEDULE_DATAFLOW = 4          /* one co-resident launch, NO barriers: vertex records carry the stage that wrote them and
	                                     every element re-gathers until its four records carry the expected stage.  Same serial
	                                     order and bits as the others, also with volume passes and post-solve damping sweeps
	                                     (k_substeps_dataflow_general).  Only calls with in-constraint Rayleigh damping
	                                     (Paper / Limit) run as XF_SCHEDULE_PERSISTENT. */
} xf_schedule;

/* Clustered colouring: consecutive stream elements that share vertices and touch at most 8 distinct ones (MeshGen: the six
 * tets of a cell) form a cluster solved back to back by ONE thread on the barrier-free schedule; clusters are coloured
 * instead of elements (8 classes on the lattice instead of 24).  For every schedule the result is an ordinary colouring
 * (colour = clusterColour * groupSize + position in cluster) and xf_get_order reports the equivalent serial order. */
typedef enum xf_grouping {
	XF_GROUPING_AUTO = 0,     /* = XF_GROUPING_ELEMENTS (clusters measured slower at 1M tets, faster at 2M) */
	XF_GROUPING_ELEMENTS = 1, /* colour single elements (colorHint honoured) */
	XF_GROUPING_CLUSTERS = 2, /* clusters, whatever the schedule (falls back to elements if that needs > 254 colours) */
	/* Elements coloured singly (as XF_GROUPING_ELEMENTS: same colouring, same serial order, same bits), but on the
	 * barrier-free schedule the thread that ran position j of one colour keeps, in private shared-memory slots, the
	 * vertex records that position j of the next colour uses again when nobody writes them in between: those corners
	 * cost no L2 gather / scatter.  With xf_generate_tet_block's ring-ordered hint a cell's six tets move 18 records
	 * through L2 instead of 48.  Applies when every colour fits one wave of the co-resident grid (else: ELEMENTS). */
	XF_GROUPING_CHAINS = 3
} xf_grouping;

typedef struct xf_create_params {
	uint32_t abiVersion;   /* XF_ABI_VERSION */
	int32_t device;        /* CUDA device ordinal */
	float density;         /* Sim::AddBlock passes 1.0f, 2.0f for the Armadillo (Demo.cpp:123) */
	int32_t autoResize;    /* GeoLinear3d::Init's autoResize (Geo.cpp:726-728) */
	int32_t precision;     /* xf_precision */
	int32_t schedule;      /* xf_schedule */
	void* stream;          /* cudaStream_t to enqueue on, or NULL for a library-owned stream */
	const uint32_t* colorHint; /* optional: colour per element (in stream order); validated, never trusted */
	uint32_t colorHintCount;   /* number of entries in colorHint (must equal the element count) */
	uint32_t grouping;         /* xf_grouping */
	uint32_t partition;      /* xf_part_create only: xf_partition_method */
	uint32_t _reserved[3];
} xf_create_params;

typedef struct xf_scene xf_scene;

const char* xf_last_error(void);
int xf_device_count(int* outCount);
void xf_default_create_params(xf_create_params* p);

/* ---- mesh producer (host only; MeshGen.cpp:156-244) ----
 * nodes: 3*(w+1)(h+1)(d+1) floats; idxStream: 30*w*h*d u32 in the reference's stream format
 * [4, v0, v1, v2, v3]*; colorHint (optional, may be NULL): 6*w*h*d entries, an analytic
 * 24-colouring for XF_PATTERN_UNIFORM (0xffffffff everywhere for other patterns). */
int xf_generate_tet_block(uint32_t width, uint32_t height, uint32_t depth, float sx, float sy, float sz, uint32_t pattern,
                          float wonkiness, float* nodes, uint32_t* idxStream, uint32_t* colorHint);

/* ---- scene lifetime (GeoLinear3d::Init, Geo.cpp:697-772) ---- */
int xf_create(const xf_create_params* params, const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream,
              uint32_t idxCount, xf_scene** outScene);
int xf_destroy(xf_scene* scene);

uint32_t xf_vert_count(const xf_scene* scene);
uint32_t xf_element_count(const xf_scene* scene);
uint32_t xf_color_count(const xf_scene* scene);
/* The serial element order the device schedule is equivalent to (inject into GeoLinear3d::tOrder). */
int xf_get_order(const xf_scene* scene, uint32_t* order);
int xf_get_colors(const xf_scene* scene, uint32_t* colorOfElement);
/* Per-element constants as InitFiniteElement produced them (Fem.cpp:196-224); any pointer may be NULL. */
/* Barrier-free schedule (XF_SCHEDULE_DATAFLOW): for the k-th element of xf_get_order, the stage code of the previous writer of
 * each of its corners (0 = the substep's vertex phase, else 1 + colour); per vertex the code of its last writer (0 = none). */
int xf_get_stage_codes(const xf_scene* scene, uint8_t* predCode4, uint8_t* lastCode);
/* XF_GROUPING_CHAINS: for the k-th element of xf_get_order one word, per corner n bits [5n, 5n+5) = private slot (0-3) |
 * 8 if the record is gathered from L2 (waiting for its stage tag) | 16 if it is scattered to L2 after the solve; all zero when
 * the scene is not chained.  outPermille = corner uses served from a slot, per 1000.  Either pointer may be NULL. */
int xf_get_chain_info(const xf_scene* scene, uint32_t* info, uint32_t* outPermille);
/* Codes of the count-versioned velocity records behind the barrier-free damping sweeps: rank4 = 4 per element
 * in the serial order of xf_get_order, below8 = 8 per vertex (elements around it below each eighth of the serial order). */
int xf_get_damping_codes(const xf_scene* scene, uint8_t* rank4, uint8_t* below8);
int xf_get_elements(const xf_scene* scene, uint32_t* idx4, float* Qi9, float* QQ3, float* QR3, float* volume, float* surfaceArea);

/* ---- stepping (Geo3d::Substep, Geo.cpp:305-356), n substeps with tickId advancing per substep ---- */
int xf_substep(xf_scene* scene, const xf_settings* settings, const xf_manipulator* manip, float dt, uint32_t n);
/* n substeps in ONE launch while Settings::lockedRightTransform3d and / or Manipulator::pickDirTarget change from substep to
 * substep - what Sim::Update does between its Geo::Substep calls (Demo.cpp:67-90).  lockT3d: n x 12 floats or NULL;
 * pickDirTarget: n x 3 floats or NULL (used when manip->picked).  tickId advances by one per substep.  Bit-identical to n calls
 * of xf_substep(.., 1) with those values.  xf_frame_update uses it, so an interactive frame (dragging, animated lock) is one
 * launch. */
int xf_substep_varying(xf_scene* scene, const xf_settings* settings, const xf_manipulator* manip, float dt, uint32_t n, const float* lockT3d,
                       const float* pickDirTarget);
int xf_sync(xf_scene* scene);

/* Extensions named by the task that the reference lacks (semantics in DESIGN.md §Extensions). */
int xf_set_ground(xf_scene* scene, int enabled, float y0, float friction);
int xf_set_handles(xf_scene* scene, uint32_t count, const uint32_t* vertIdx, const float* targetXYZ);

/* ---- state (host buffers, packed xyz triples of doubles like dvec3 without its padding) ---- */
int xf_get_state(xf_scene* scene, double* X, double* V, float* w);
int xf_set_state(xf_scene* scene, const double* X, const double* V, const float* w);
int xf_get_rest(xf_scene* scene, double* X0, double* O, uint8_t* flags);
int xf_get_origin(const xf_scene* scene, float* origin3);
/* Asynchronous variants for pinned host buffers (e2e loop): enqueue on the scene's stream, no sync. */
int xf_get_state_async(xf_scene* scene, double* X, double* V);
int xf_set_state_async(xf_scene* scene, const double* X, const double* V);

int xf_transform(xf_scene* scene, const float* m9 /* column-major mat3 */);
/* fp32 sum in element order, bit-identical to GeoLinear3d::CalculateVolume (per-element terms on the device). */
int xf_volume(xf_scene* scene, float* outVolume);

/* Device-side statistics (fp64 reductions): [0] volume, [1] kinetic energy, [2] gravitational potential,
 * [3] deviatoric elastic energy, [4] volumetric elastic energy, [5] count of non-finite position components. */
int xf_stats(xf_scene* scene, const xf_settings* settings, double* out6);

/* Introspection for benchmarks. */
typedef struct xf_info {
	uint32_t vertCount, elementCount, colorCount, minColorSize, maxColorSize;
	uint32_t smCount, gridBlocks, blockThreads;
	uint32_t elementRecordBytes; /* bytes streamed per element per sweep in the active precision */
	uint32_t schedule;           /* resolved xf_schedule */
	uint64_t kernelLaunches;     /* kernels launched by this scene so far */
	uint64_t l2Bytes;            /* cudaDeviceProp::l2CacheSize */
	uint32_t chainedPermille;    /* XF_GROUPING_CHAINS: corner uses served from the thread's private slots, per 1000 (else 0) */
	uint32_t lastKernel;         /* xf_kernel_id of the stepping kernel the last xf_substep launched (0 = none yet) */
} xf_info;
/* which stepping kernel ran (xf_info::lastKernel): lets a caller (and the tests) see when a call left the barrier-free path */
typedef enum xf_kernel_id {
	XF_KERNEL_NONE = 0,
	XF_KERNEL_DATAFLOW = 1,      /* k_substeps_dataflow: barrier-free, versioned records */
	XF_KERNEL_CHAIN = 2,         /* k_substeps_chain */
	XF_KERNEL_CLUSTER = 3,       /* k_substeps_cluster */
	XF_KERNEL_PERSISTENT = 4,    /* k_substeps_persistent: grid barrier per colour */
	XF_KERNEL_BRICKS = 5,        /* retired (never reported) */
	XF_KERNEL_PER_COLOR = 6,     /* k_sweep_color / k_vertex_phase, one launch per colour */
	XF_KERNEL_DATAFLOW_GENERAL = 7 /* k_substeps_dataflow_general: barrier-free with volume passes / post-solve damping sweeps */
} xf_kernel_id;
int xf_get_info(const xf_scene* scene, xf_info* out);

/* ---- frame driver: Sim::Update (Demo.cpp:37-103) for one Geo ----
 * Per frame: substep count from the wall-clock dt (clamped to ceil(medianFrameTime / sdt)), the time-corrected
 * PBD-damping / drag constants written into *settings like the reference does, the right-side lock transform and the
 * manipulator ray animated with substep resolution, tickId advanced.  xf_frame_state holds Sim's persistent scalars
 * (Demo.h:40-44). */
typedef struct xf_frame_state {
	float dtResidual;
	uint32_t tickId;
	float leftRightSeparationOld;
	float rightRotationTheta;
} xf_frame_state;
void xf_frame_state_init(xf_frame_state* state);
int xf_frame_update(xf_scene* scene, xf_settings* settings, xf_manipulator* manip, float dt, float medianFrameTime, xf_frame_state* state,
                    uint32_t* outSubsteps);

/* ---- Sim (Demo.h:18-52): several Geos under one Settings block and one frame clock ----
 * xf_sim_add_block = Sim::AddBlock (Demo.cpp:120-154; density 2 when autoResize, Demo.cpp:123), xf_sim_add_block_from_settings
 * = the block Demo::UpdateSettings builds from the shape / pattern / wonkiness in the Settings (table Demo.cpp:289-318,
 * exposed by xf_block_from_settings), xf_sim_finish_adding_blocks = Sim::FinishAddingBlocks (Demo.cpp:156-168),
 * xf_sim_set_geo_offset = Sim::SetGeoOffset (Demo.cpp:179-186), xf_sim_update = Sim::Update (Demo.cpp:37-103) with the
 * manipulator acting on geo `pickedGeo` (Manipulator::pickedGeo; -1 = none), xf_sim_reset = Sim::Reset.  Only Element_T4
 * blocks; every geo is an xf_scene (xf_sim_geo) on params->device / stream and is stepped with one launch per frame. */
typedef struct xf_sim xf_sim;
int xf_sim_create(const xf_create_params* params, xf_sim** outSim);
int xf_sim_destroy(xf_sim* sim);
int xf_sim_reset(xf_sim* sim);
int xf_sim_add_block(xf_sim* sim, uint32_t elementType, const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream,
                     uint32_t idxCount, int autoResize, const uint32_t* colorHint, uint32_t colorHintCount);
int xf_block_from_settings(const xf_settings* settings, uint32_t* outWidth, uint32_t* outHeight, float* outScaleX, float* outScaleY,
                           uint32_t* outPattern);
int xf_sim_add_block_from_settings(xf_sim* sim, const xf_settings* settings);
int xf_sim_finish_adding_blocks(xf_sim* sim, const xf_settings* settings);
int xf_sim_set_geo_offset(xf_sim* sim, float x, float y);
int xf_sim_update(xf_sim* sim, xf_settings* settings, xf_manipulator* manip, int pickedGeo, float dt, float medianFrameTime,
                  uint32_t* outSubsteps);
uint32_t xf_sim_geo_count(const xf_sim* sim);
xf_scene* xf_sim_geo(xf_sim* sim, uint32_t index);
float xf_sim_volume0(const xf_sim* sim, uint32_t index);
int xf_sim_get_frame_state(const xf_sim* sim, xf_frame_state* out);

/* ---- batched scenes (BASELINE config 3): nScenes independent instances of ONE rest mesh, each with its own
 * state and Settings, e.g. a vector of RL environments.  The reference would hold them as nScenes Geo objects and
 * call Geo::Substep on each (Sim::Update's `for geo` loop, Demo.cpp:86-88); here one call steps all of them, one
 * thread group per scene with the scene resident in shared memory.  Scenes too large for shared memory are
 * rejected with XF_ERR_UNSUPPORTED (use xf_create per scene).  State buffers are [scene][vertex][xyz]. ---- */
typedef struct xf_batch xf_batch;
int xf_batch_create(const xf_create_params* params, const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream,
                    uint32_t idxCount, uint32_t nScenes, xf_batch** outBatch);
int xf_batch_destroy(xf_batch* batch);
uint32_t xf_batch_scene_count(const xf_batch* batch);
uint32_t xf_batch_vert_count(const xf_batch* batch);
uint32_t xf_batch_element_count(const xf_batch* batch);
uint32_t xf_batch_color_count(const xf_batch* batch);
int xf_batch_get_order(const xf_batch* batch, uint32_t* order);
int xf_batch_set_ground(xf_batch* batch, int enabled, float y0, float friction);
/* settingsCount == 1 (shared) or == scene count.  Energy / solve mode / Rayleigh type must agree across scenes. */
int xf_batch_substep(xf_batch* batch, const xf_settings* settings, uint32_t settingsCount, float dt, uint32_t n);
int xf_batch_sync(xf_batch* batch);
int xf_batch_get_state(xf_batch* batch, uint32_t firstScene, uint32_t count, double* X, double* V, float* w);
int xf_batch_set_state(xf_batch* batch, uint32_t firstScene, uint32_t count, const double* X, const double* V, const float* w);
int xf_batch_get_info(const xf_batch* batch, uint32_t* groupThreads, uint32_t* blockThreads, uint32_t* smemBytes, uint64_t* launches);

/* ---- one mesh on several GPUs (BASELINE config 4) ----
 * Every rank (one process per GPU) calls xf_part_create with the SAME full mesh and its own rank; the library
 * computes the same partition plan everywhere (x-slabs of elements, local copies of the touched vertices, global
 * colouring).  The per-colour halo exchange is fused into the sweep kernels as stores into the peers' vertex
 * arrays over NVLink (CUDA IPC mappings) plus flag signalling; the host only has to move each rank's 128-byte
 * xf_part_ipc_export blob to all ranks once (torch.distributed / MPI / files) and hand them to
 * xf_part_ipc_connect.  All ranks must then issue the same xf_part_substep calls.  Results equal the
 * single-GPU schedule bit for bit.  device < 0 creates a host-only plan (no stepping) whose halo lists can be
 * inspected.  Calls with damping (in-constraint or post-solve sweeps) or volume passes run on the flag protocol, one launch per
 * phase, with the velocities of shared vertices mirrored like the positions; the plain sweep on the schedule chosen at creation. */
typedef struct xf_partition xf_partition;
#define XF_IPC_BYTES 128
int xf_part_create(const xf_create_params* params, const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream,
                   uint32_t idxCount, uint32_t nRanks, uint32_t rank, xf_partition** outPart);
int xf_part_destroy(xf_partition* part);
uint32_t xf_part_local_vert_count(const xf_partition* part);
uint32_t xf_part_local_element_count(const xf_partition* part);
uint32_t xf_part_peer_count(const xf_partition* part);
uint32_t xf_part_color_count(const xf_partition* part);
uint32_t xf_part_global_vert_count(const xf_partition* part);
uint32_t xf_part_global_element_count(const xf_partition* part);
int xf_part_get_local_verts(const xf_partition* part, uint32_t* localToGlobal);
int xf_part_get_local_elements(const xf_partition* part, uint32_t* globalElementIds, uint32_t* colorStart /* colours + 1 */);
int xf_part_get_peers(const xf_partition* part, uint32_t* peerRanks);
/* vertices (local ids, ascending global id) sent to / received from peer `peerSlot` after colour `color` */
int xf_part_get_halo(const xf_partition* part, uint32_t color, uint32_t peerSlot, int send, uint32_t* count, uint32_t* localVerts);
int xf_part_get_order(const xf_partition* part, uint32_t* fullOrder);
int xf_part_get_global_color_start(const xf_partition* part, uint32_t* colorStart);
int xf_part_get_initial(const xf_partition* part, float* w, uint8_t* flags);
/* Barrier-free schedule (XF_SCHEDULE_DATAFLOW; AUTO picks it when every vertex has at most two copies): per local element
 * 4 codes naming the previous writer of each corner in the global serial order (0 = the substep's vertex phase, 255 = the
 * same for a shared vertex, else 1 + colour), per local vertex the code of its last writer. */
int xf_part_get_dataflow_codes(const xf_partition* part, uint8_t* predCode4, uint8_t* lastCode, int* outOk);
int xf_part_ipc_export(xf_partition* part, void* out128);
int xf_part_ipc_connect(xf_partition* part, const void* allRanksBlobs);
int xf_part_set_ground(xf_partition* part, int enabled, float y0, float friction);
int xf_part_substep(xf_partition* part, const xf_settings* settings, float dt, uint32_t n);
int xf_part_sync(xf_partition* part);
int xf_part_get_state(xf_partition* part, double* X, double* V, float* w); /* local vertices */
int xf_part_get_info(const xf_partition* part, uint64_t* launches, uint64_t* epoch);

/* ---- measurement / test hooks (no counterpart in the reference; used by bench.py, tools/ and tests/) ----
 * xf_debug_l2_bandwidth: streaming copy (mode 0: read + write bytes per second) or read (mode 1) over buffers of `bytes`
 *   each that stay resident in L2, 256-bit accesses, best of `reps` launches of `passes` passes: the L2 peak the roofline
 *   of an L2-resident mesh is quoted against.
 * xf_debug_torn_records: stress of the hardware property the barrier-free schedules rely on - a 32-byte-aligned 256-bit
 *   access is one transaction.  Writer CTAs rewrite records whose four 8-byte words encode one counter, reader CTAs count
 *   records with mixed words.  remoteDevice >= 0: the writers run on `device` and store over NVLink into `remoteDevice`'s
 *   memory while `remoteDevice` reads locally (the partitioned schedule's pattern).
 * xf_debug_scene_knob: 0 = first stage tag of the next barrier-free launch (tag wrap-around test), 1 = polls before a
 *   waiting warp gives up, 2 = corrupt the last-writer code of device vertex 0 (stall-report test).
 * xf_debug_barrier_us: cost of a bare grid barrier (several implementations). */
int xf_debug_l2_bandwidth(int device, int mode, uint64_t bytes, uint32_t passes, int reps, int blocksPerSm, double* outGBs);
/* cycles of (mode 0) one element solve of a lone warp, (mode 1) one record hand-off between two SMs, (mode 2) one dependent
 * 256-bit L2 load: the decomposition of a stage of the barrier-free sweep (xf_probe_latency.cu) */
int xf_debug_stage_latency(int device, int mode, uint32_t iterations, double* outCycles);
/* The warp-cooperative element solve (FOUR lanes per element, xf_element_coop.cuh) against the one-thread solve of the stepping
 * kernels on the same gathered elements (MixedSel / YeohSkinFast, simultaneous, undamped): elemConsts = nElems x 16 floats {Qi[9]
 * ([col][row]), volume, QQ[3], QR[3]} as xf_get_elements returns them, X = nElems x 12 doubles, w = nElems x 4, params4 = {1 + mu/lambda,
 * 1/mu, 1/lambda, dt^2}.  out8 = {cycles per solve of a lone warp: one-thread, four-lane; element solves per second with warpsPerSm
 * warps on every SM: one-thread, four-lane; doubles that differ between the variants after `iterations` chained solves, doubles
 * compared; SM count, SM clock in kHz}; outXSingle / outXCoop (nElems x 12 doubles, or NULL) receive the final positions.
 * variant bit 0: 0 = lane 3 broadcasts vertex 3's position, 1 = every lane gathered it itself (one shuffle stage less; the probe
 * gathers once, so the two sides are comparable bit for bit for iterations == 1 only); bit 1: the one-thread side runs the scalar
 * arithmetic instead of the two-wide one (same bits). */
int xf_debug_coop_element(int device, int energy, int variant, const float* elemConsts, const double* X, const float* w, uint32_t nElems,
                          const float* params4, uint32_t iterations, int warpsPerSm, double* out8, double* outXSingle, double* outXCoop);
int xf_debug_torn_records(int device, int remoteDevice, uint32_t nRecords, uint32_t rounds, uint64_t* outReads, uint64_t* outTorn);
int xf_debug_scene_knob(xf_scene* scene, int knob, uint32_t value);
int xf_debug_barrier_us(int device, int variant, int blocksPerSm, int threads, uint32_t iterations, float* outUsPerBarrier);

#ifdef __cplusplus
}
#endif
#endif /* XPBD_FEM_B200_H */
