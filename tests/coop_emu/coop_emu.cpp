// Host emulation of the warp-cooperative element solve (xpbd-fem_b200/csrc/xf_element_coop.cuh): FOUR THREADS play the four
// lanes of a quad, shuffles and the quad's shared-memory row go through a mailbox with a barrier on both sides, arithmetic is
// plain IEEE fp32 / fp64 (this file is compiled with -ffp-contract=off, so a*b+c is never fused - what __fmul_rn / __fadd_rn
// guarantee on the device).  The header is the SAME source the device compiles; only the lane policy differs.
// Test infrastructure (tests/test_coop_emu.py): lets the lane assignment and the association order of every scalar be checked
// bit for bit against the reference on a machine without a GPU.
#define XF_COOP_FN inline
#include "../../include/xpbd_fem_b200.h"
#include "../../xpbd-fem_b200/csrc/xf_element_coop.cuh"

#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct SpinBarrier {
	std::atomic<int> count{ 0 };
	std::atomic<int> phase{ 0 };
	void wait() {
		const int ph = phase.load(std::memory_order_acquire);
		if (count.fetch_add(1, std::memory_order_acq_rel) == 3) {
			count.store(0, std::memory_order_relaxed);
			phase.store(ph + 1, std::memory_order_release);
		} else {
			int spins = 0;
			while (phase.load(std::memory_order_acquire) == ph) {
				if (++spins > 200) { std::this_thread::yield(); }
			}
		}
	}
};

struct Quad {
	SpinBarrier bar;
	double boxD[4];
	float boxF[4];
	float sm[xf::kCoopQuadFloats];
};

struct HostOps {
	static float mul(float a, float b) { return a * b; }
	static float add(float a, float b) { return a + b; }
	static float sub(float a, float b) { return a - b; }
	static float rcp(float x) { return 1.0f / x; }
	static float max(float a, float b) { return a > b ? a : b; } // fmaxf for non-NaN operands
	static float sel(bool cond, float a, float b) { return cond ? a : b; }
	static double dsub(double a, double b) { return a - b; }
	static double dadd(double a, double b) { return a + b; }
	static float d2f(double a) { return (float)a; }
	static double f2d(float a) { return (double)a; }
};

struct HostLane4 {
	typedef HostOps O;
	Quad* quad;
	int lane;
	int q() const { return lane; }
	float shfl(float v, int src) const {
		quad->boxF[lane] = v;
		quad->bar.wait();
		const float r = quad->boxF[src];
		quad->bar.wait();
		return r;
	}
	double shfl(double v, int src) const {
		quad->boxD[lane] = v;
		quad->bar.wait();
		const double r = quad->boxD[src];
		quad->bar.wait();
		return r;
	}
	void sts(int i, float v) const { quad->sm[i] = v; }
	void lds3(int i, float (&out)[3]) const { out[0] = quad->sm[i]; out[1] = quad->sm[i + 1]; out[2] = quad->sm[i + 2]; }
	void sync() const { quad->bar.wait(); }
};

struct Elem {
	float Qi[3][3];
	float QQ[3], QR[3];
};

}  // namespace

// One Gauss-Seidel sweep over `order` (element ids) of the mesh {idx4, element constants}, every element solved by the four-lane
// code; X (3 doubles per vertex) is updated in place.  energy: XF_ENERGY_MIXED_SEL or XF_ENERGY_YEOH_SKIN_FAST; x3Gathered: every
// lane reads vertex 3 itself instead of receiving it from lane 3.
extern "C" int coop_emu_sweep(int energy, int x3Gathered, const uint32_t* idx4, const float* Qi9, const float* QQ3, const float* QR3, const float* volume,
                              float a, float invMu, float invLambda, float dt2, double* X, const float* w, const uint32_t* order,
                              uint32_t nOrder) {
	if (energy != XF_ENERGY_MIXED_SEL && energy != XF_ENERGY_YEOH_SKIN_FAST) { return 1; }
	Quad quad;
	memset(quad.sm, 0, sizeof(quad.sm));
	auto lane = [&](int l) {
		HostLane4 ln{ &quad, l };
		for (uint32_t k = 0; k < nOrder; k++) {
			const uint32_t t = order[k];
			Elem e;
			memcpy(e.Qi, Qi9 + 9 * (size_t)t, sizeof(e.Qi));
			memcpy(e.QQ, QQ3 + 3 * (size_t)t, sizeof(e.QQ));
			memcpy(e.QR, QR3 + 3 * (size_t)t, sizeof(e.QR));
			// comp = {1/mu/vol, 1/lambda/vol}, alpha = comp / dt^2 (Fem.cpp:449, Xpbd.h:154); 0 / y == +0 for y > 0
			const float vol = volume[t];
			const float comp0 = invMu / vol;
			const float comp1 = (invLambda == 0.0f && vol > 0.0f) ? 0.0f : invLambda / vol;
			const float alpha0 = comp0 / dt2;
			const float alpha1 = (comp1 == 0.0f && dt2 > 0.0f) ? 0.0f : comp1 / dt2;
			const uint32_t v = idx4[4 * (size_t)t + l];
			double x[3] = { X[3 * (size_t)v], X[3 * (size_t)v + 1], X[3 * (size_t)v + 2] };
			const uint32_t v3 = idx4[4 * (size_t)t + 3];
			const double x3[3] = { X[3 * (size_t)v3], X[3 * (size_t)v3 + 1], X[3 * (size_t)v3 + 2] };
			quad.bar.wait(); // every lane has read vertex 3 before lane 3 rewrites it
			if (energy == XF_ENERGY_YEOH_SKIN_FAST) {
				if (x3Gathered) {
					xf::SolvePrefactoredSimulCoop4<XF_ENERGY_YEOH_SKIN_FAST, true>(ln, a, e, alpha0, alpha1, x, w[v], x3);
				} else {
					xf::SolvePrefactoredSimulCoop4<XF_ENERGY_YEOH_SKIN_FAST, false>(ln, a, e, alpha0, alpha1, x, w[v], x3);
				}
			} else {
				if (x3Gathered) {
					xf::SolvePrefactoredSimulCoop4<XF_ENERGY_MIXED_SEL, true>(ln, a, e, alpha0, alpha1, x, w[v], x3);
				} else {
					xf::SolvePrefactoredSimulCoop4<XF_ENERGY_MIXED_SEL, false>(ln, a, e, alpha0, alpha1, x, w[v], x3);
				}
			}
			X[3 * (size_t)v] = x[0]; X[3 * (size_t)v + 1] = x[1]; X[3 * (size_t)v + 2] = x[2];
			quad.bar.wait(); // the next element gathers what this one scattered
		}
	};
	std::vector<std::thread> th;
	for (int l = 0; l < 4; l++) { th.emplace_back(lane, l); }
	for (auto& t : th) { t.join(); }
	return 0;
}
