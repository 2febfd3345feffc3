// Runtime (energy, simultaneous, exact, damped) -> template instantiation, shared by xf_kernels.cu and xf_batch.cu.
#pragma once

#include <atomic>

#include "xf_scene.h"

namespace xf {

// Resident CTAs per SM of one kernel instantiation, cached per DEVICE (a process may hold scenes on several GPUs, and the
// shared-memory opt-in is a per-device function attribute).  Lock-free: racing host threads compute the same value.
struct OccupancyCache {
	static constexpr int kMaxDevices = 64;
	std::atomic<int> perSm[kMaxDevices];
	OccupancyCache() { for (auto& v : perSm) { v.store(0, std::memory_order_relaxed); } }
	cudaError_t Get(const void* fn, int threads, size_t smem, bool optInSmem, int* out) {
		int dev = 0;
		cudaError_t e = cudaGetDevice(&dev);
		if (e != cudaSuccess) { return e; }
		const bool cached = dev >= 0 && dev < kMaxDevices;
		int v = cached ? perSm[dev].load(std::memory_order_acquire) : 0;
		if (v == 0) {
			if (optInSmem) {
				e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
				if (e != cudaSuccess) { return e; }
			}
			e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, fn, threads, smem);
			if (e != cudaSuccess) { return e; }
			if (v < 1) { return cudaErrorLaunchOutOfResources; }
			if (cached) { perSm[dev].store(v, std::memory_order_release); }
		}
		*out = v;
		return cudaSuccess;
	}
};

template <template <int, bool, bool, bool> class Fn, typename... Args>
cudaError_t DispatchConfig(uint32_t energy, bool simul, bool exact, bool damped, Args&&... args) {
#define XF_CASE(E)                                                                                                      \
	case E:                                                                                                             \
		if (simul) {                                                                                                    \
			if (exact) { return damped ? Fn<E, true, true, true>::Run(args...) : Fn<E, true, true, false>::Run(args...); } \
			return damped ? Fn<E, true, false, true>::Run(args...) : Fn<E, true, false, false>::Run(args...);           \
		}                                                                                                               \
		if (exact) { return damped ? Fn<E, false, true, true>::Run(args...) : Fn<E, false, true, false>::Run(args...); } \
		return damped ? Fn<E, false, false, true>::Run(args...) : Fn<E, false, false, false>::Run(args...);
	switch (energy) {
		XF_CASE(XF_ENERGY_MIXED)
		XF_CASE(XF_ENERGY_MIXED_SEL)
		XF_CASE(XF_ENERGY_YEOH_SKIN)
		XF_CASE(XF_ENERGY_YEOH_SKIN_FAST)
	default: return cudaErrorInvalidValue;
	}
#undef XF_CASE
}


}  // namespace xf
