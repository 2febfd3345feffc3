// Frame driver: the reference's Sim::Update (Demo.cpp:37-103) for one Geo, on top of xf_substep.
// Decides the substep count from the frame's wall-clock dt, derives the time-corrected damping / drag constants,
// animates the right-side lock and the manipulator ray with substep resolution, and advances tickId.
// Host-only arithmetic, compiled with -ffp-contract=off so every derived constant matches the reference's bits.
#include <cmath>
#include <cstring>
#include <vector>

#include "xf_scene.h"

extern "C" void xf_frame_state_init(xf_frame_state* st) { // Sim::Reset, Demo.cpp:170-177
	if (!st) { return; }
	st->dtResidual = 0.0f;
	st->tickId = 0;
	st->leftRightSeparationOld = 1.0f;
	st->rightRotationTheta = 0.0f;
}

// Sim::Update for `count` Geos that share the Settings (Demo.cpp:86-88: every substep steps every geo; here every geo gets its
// frame's substeps in one launch - the geos are independent, so the order of (geo, substep) pairs does not matter).  The
// manipulator acts on geo `pickedGeo` only (Manipulator::pickedGeo, Geo.cpp:334).
int xf::FrameUpdateGeos(xf_scene* const* scenes, uint32_t count, int pickedGeo, xf_settings* settings, xf_manipulator* manip, float dt,
                        float medianFrameTime, xf_frame_state* st, uint32_t* outSubsteps) {
	// count == 0: bookkeeping only (substep count, derived constants, lock / manipulator animation) - lets the
	// host logic be tested without a device; nothing is stepped.
	if (!settings || !st) { return xf::Fail(XF_ERR_INVALID, "null argument"); }
	const float kSpacing = (20.0f / 100.0f) / (float)(31); // Demo.cpp:16
	const uint32_t kRotate90 = 1u << 24, kRotateLock = 1u << 28; // Settings.h:24, 28

	// how many substeps to take, Demo.cpp:39-49
	const float sdt = 1.0f / settings->substepsPerSecond;
	st->dtResidual += settings->timeScale * dt;
	uint32_t substeps = (uint32_t)(st->dtResidual / sdt);
	const uint32_t maxSubsteps = (uint32_t)ceilf((float)medianFrameTime / sdt);
	if (substeps > maxSubsteps) {
		substeps = maxSubsteps;
		st->dtResidual = 0.0f;
	} else {
		st->dtResidual -= sdt * (float)substeps;
	}

	// frame-dependent settings, Demo.cpp:51-63
	const float tcPbd = 1.0f - powf(1.0f - settings->pbdDamping, 1000.0f * sdt);
	settings->areaAndTimeCorrectedPbdDamping = tcPbd * kSpacing * kSpacing;
	settings->volumeAndTimeCorrectedPbdDamping = tcPbd * 6.0f * kSpacing * kSpacing;
	const float amPbd = 1.0f - powf(1.0f - settings->pbdDamping, 1000.0f * (float)XF_AMORTIZATION_PERIOD * sdt);
	settings->amortizedAreaAndTimeCorrectedPbdDamping = amPbd * kSpacing * kSpacing;
	settings->amortizedVolumeAndTimeCorrectedPbdDamping = amPbd * 6.0f * kSpacing * kSpacing;
	settings->timeCorrectedDrag = 1.0f - powf(1.0f - settings->drag, 1000.0f * sdt);

	const bool lockRight = (settings->flags & XF_SETTINGS_LOCK_RIGHT) != 0;
	const bool picked = manip && manip->picked;
	if (outSubsteps) { *outSubsteps = substeps; }

	if (!lockRight && !picked) {
		// nothing varies inside the frame except tickId (advanced by the kernels): one launch for all substeps
		settings->tickId = st->tickId;
		if (manip && substeps > 0) { // the reference still lerps the ray every substep; its final value is pickDir
			const float a = 1.0f;
			for (int k = 0; k < 3; k++) { manip->pickDirTarget[k] = manip->pickDirOld[k] * (1.0f - a) + manip->pickDir[k] * a; }
		}
		if (substeps > 0) {
			for (uint32_t g = 0; g < count; g++) {
				int rc = xf_substep(scenes[g], settings, (int)g == pickedGeo ? manip : nullptr, sdt, substeps);
				if (rc != XF_OK) { return rc; }
			}
			st->tickId += substeps;
			settings->tickId = st->tickId - 1; // the value the reference leaves in its Settings after the loop
		}
	} else if (substeps > 0) {
		// the lock transform and the manipulator ray change every substep: their per-substep values go to the device as arrays and
		// the whole frame is still ONE launch (xf_substep_varying)
		std::vector<float> lockRows(lockRight ? 12 * (size_t)substeps : 0), dirRows(manip ? 3 * (size_t)substeps : 0);
		const uint32_t firstTick = st->tickId;
		for (uint32_t substep = 0; substep < substeps; substep++) {
			if (lockRight) { // Demo.cpp:69-80
				const float a = (float)substep / (float)substeps;
				const float lrs = st->leftRightSeparationOld * (1.0f - a) + settings->leftRightSeparation * a;
				const float sx = lrs * 2.0f - 1.0f;
				if (settings->flags & kRotateLock) {
					st->rightRotationTheta += sdt;
					if (st->rightRotationTheta > (float)(2.0 * M_PI)) { st->rightRotationTheta -= (float)(2.0 * M_PI); }
				} else {
					st->rightRotationTheta = (settings->flags & kRotate90) ? (float)(1.5f * M_PI) : 0.0f;
				}
				const float c = cosf(st->rightRotationTheta), s = sinf(st->rightRotationTheta);
				// rotation(theta) * scale with rotation = [(c, s), (-s, c)], scale = [(sx, 0), (0, 1)]; mat2*vec2 = row dots
				float* T = settings->lockedRightTransform;
				T[0] = c * sx + (-s) * 0.0f;
				T[1] = s * sx + c * 0.0f;
				T[2] = c * 0.0f + (-s) * 1.0f;
				T[3] = s * 0.0f + c * 1.0f;
				float* T3 = settings->lockedRightTransform3d;
				T3[0] = T[0]; T3[1] = T[1]; T3[2] = 0.0f;
				T3[4] = T[2]; T3[5] = T[3]; T3[6] = 0.0f;
				T3[8] = 0.0f; T3[9] = 0.0f; T3[10] = 1.0f;
				memcpy(&lockRows[12 * (size_t)substep], T3, sizeof(float) * 12);
			}
			if (manip) { // Demo.cpp:84
				const float a = (float)(substep + 1) / (float)substeps;
				for (int k = 0; k < 3; k++) { manip->pickDirTarget[k] = manip->pickDirOld[k] * (1.0f - a) + manip->pickDir[k] * a; }
				memcpy(&dirRows[3 * (size_t)substep], manip->pickDirTarget, sizeof(float) * 3);
			}
			++st->tickId;
		}
		settings->tickId = firstTick;
		for (uint32_t g = 0; g < count; g++) {
			const bool mine = (int)g == pickedGeo && manip;
			int rc = xf_substep_varying(scenes[g], settings, mine ? manip : nullptr, sdt, substeps, lockRight ? lockRows.data() : nullptr,
			                            mine ? dirRows.data() : nullptr);
			if (rc != XF_OK) { return rc; }
		}
		settings->tickId = st->tickId - 1; // the value the reference leaves in its Settings after the loop
	}
	st->leftRightSeparationOld = settings->leftRightSeparation;
	return XF_OK;
}

extern "C" int xf_frame_update(xf_scene* scene, xf_settings* settings, xf_manipulator* manip, float dt, float medianFrameTime, xf_frame_state* st,
                               uint32_t* outSubsteps) {
	return xf::FrameUpdateGeos(&scene, scene ? 1u : 0u, 0, settings, manip, dt, medianFrameTime, st, outSubsteps);
}
